"""CPU tests of the host-side logic around the kernels: C-ABI surface, collater indexing, AR
chunk plans, launcher environment, and the data-parallel plumbing on a 2-rank gloo group."""
import ctypes
import json
import os
import re
import socket

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# --------------------------------------------------------------------------- C ABI
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "artic.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(artic_[a-z0-9_]+)\s*\(", text)))


def test_abi_library_exports_every_declared_symbol():
    from articulatory_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 30
    assert os.path.exists(_lib.LIB_PATH), "build the library first (python __graft_entry__.py)"
    lib = ctypes.CDLL(_lib.LIB_PATH)        # loads without a GPU; no compute call is made
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/artic.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    assert lib.artic_version() >= 100
    lib.artic_arch.restype = ctypes.c_char_p
    assert lib.artic_arch() == b"sm_100a"


def test_relayout_tile_spaces_host_helpers():
    """artic_wrow_tiles / artic_wperm_tiles (host-only): every conv / linear / transposed-conv weight of the nets is taken by
    the row-run kernels; work units = groups x ceil(outer / 32) x ceil(inner tiles / chunk); a tap stride that is not the
    innermost torch index falls back to the generic tile space."""
    from articulatory_b200 import _lib
    from articulatory_b200.convspec import ConvSpec
    lib = _lib.load()

    def units(spec, chunk):
        A, B, sk, sg, sa, sb = spec.prep_strides("fwd")
        return lib.artic_wrow_tiles(spec.k, spec.groups, A, B, sk, sa, sb, chunk)

    big = ConvSpec("conv", 1024, 1024, k=5, padding=2)                      # outer = 1024 out-ch, inner = 1024 in-ch, TI = 32
    assert units(big, 1) == 32 * 32 and units(big, 8) == 32 * 4 and units(big, 100) == 32 * 4
    k41 = ConvSpec("conv", 1024, 1024, k=41, padding=20, groups=16)         # per group 64 x 64, TI = 8 (8 * 41 = 328 <= 352)
    assert units(k41, 1) == 16 * 2 * 8
    k11 = ConvSpec("conv", 128, 128, k=11, dilation=5, padding=25)          # 32 * 11 = 352 floats: TI stays 32
    assert units(k11, 1) == 4 * 4
    up = ConvSpec("convT", 256, 128, k=10, stride=5, padding=3, output_padding=1)   # outer = in-ch (torch dim 0), inner = out-ch
    assert units(up, 1) == 8 * 4
    assert units(ConvSpec("linear", 512, 256), 1) == 8 * 16
    assert units(ConvSpec("conv", 1, 128, k=15, padding=7), 1) == 4 and units(ConvSpec("conv", 32, 1, k=7, padding=3), 1) == 1
    assert lib.artic_wrow_tiles(5, 1, 64, 64, 7, 5, 320, 1) == 0            # taps not contiguous: generic tiles
    assert lib.artic_wrow_tiles(400, 1, 4, 4, 1, 400, 1600, 1) == 0         # a tap run beyond the tile's row
    assert lib.artic_wperm_tiles(5, 1, 64, 64) == 2 * 2 * 1 and lib.artic_wperm_tiles(41, 16, 64, 64) == 16 * 2 * 2 * 6


def test_scoped_planner_knobs_are_reentrant():
    """engine.weight_multicast / planner_objective set raw debug keys for the launches enqueued inside and restore the
    enclosing value on exit (nested use: the decoder's chunk capture around the engine's own scopes)."""
    from articulatory_b200 import _lib, engine
    lib = _lib.load()
    assert lib.artic_debug_get(22) == 0
    with engine.weight_multicast(2):
        assert lib.artic_debug_get(22) == 2
        with engine.weight_multicast(4):
            assert lib.artic_debug_get(22) == 4
        assert lib.artic_debug_get(22) == 2
        with engine.planner_objective(60):
            assert lib.artic_debug_get(15) == 60
        assert lib.artic_debug_get(15) == 0
    assert lib.artic_debug_get(22) == 0


def test_struct_sizes_match_header():
    """ctypes mirrors of the ABI structs have the C layout (nvcc and ctypes agree on padding)."""
    from articulatory_b200 import _lib
    assert ctypes.sizeof(_lib.Seq) == 32
    assert ctypes.sizeof(_lib.WDesc) == 8 * 8 + 7 * 8 + 12 * 4
    assert ctypes.sizeof(_lib.TapConv) % 8 == 0 and ctypes.sizeof(_lib.TapWgrad) % 8 == 0


def test_ctypes_signatures_have_the_header_arity():
    """Every prototype of include/artic.h has as many parameters as its ctypes signature (plus pointer / integer /
    float kind for each): a missing argument would shift the stream pointer into a size."""
    from articulatory_b200 import _lib
    text = open(os.path.join(ROOT, "include", "artic.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(artic_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 40
    seen = set()
    for name, args in protos:
        seen.add(name)
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        res, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, params, argtypes)
        for prm, at in zip(params, argtypes):
            is_ptr = "*" in prm
            ct_ptr = at is ctypes.c_void_p or at is ctypes.c_char_p or hasattr(at, "contents") or getattr(at, "_type_", None) is not None and isinstance(getattr(at, "_type_"), type)
            if is_ptr:
                assert ct_ptr, (name, prm, at)
            elif prm.startswith("float"):
                assert at is ctypes.c_float, (name, prm, at)
            elif prm.startswith("int64_t") or prm.startswith("long long"):
                assert ctypes.sizeof(at) == 8 and at not in (ctypes.c_double,), (name, prm, at)
            else:
                assert at in (ctypes.c_int32, ctypes.c_int), (name, prm, at)
    assert seen == set(_lib.SIGNATURES), set(_lib.SIGNATURES) ^ seen


def test_ctypes_mirrors_match_the_header_compiled_by_gcc(tmp_path):
    """include/artic.h compiled as plain C: sizeof of every ABI struct and the offset of its last field equal the ctypes
    mirrors' (a drifted mirror would hand the kernels shifted fields)."""
    import shutil
    import subprocess
    from articulatory_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = [("artic_seq_t", _lib.Seq), ("artic_tapconv_t", _lib.TapConv), ("artic_tapwgrad_t", _lib.TapWgrad),
             ("artic_resunit_t", _lib.ResUnit), ("artic_wdesc_t", _lib.WDesc), ("artic_mlp_t", _lib.Mlp),
             ("artic_disc_prep_t", _lib.DiscPrep), ("artic_adam_hyper_t", _lib.AdamHyper)]
    lines = []
    for cname, mirror in pairs:
        last = mirror._fields_[-1][0].rstrip("_") if mirror._fields_[-1][0] in ("in_",) else mirror._fields_[-1][0]
        lines.append(f'printf("%s %zu %zu\\n", "{cname}", sizeof({cname}), offsetof({cname}, {last}));')
    src = tmp_path / "sizes.c"
    src.write_text("#include <stdio.h>\n#include <stddef.h>\n#include \"artic.h\"\nint main(void) {\n" + "\n".join(lines) +
                   "\nreturn 0; }\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {ln.split()[0]: (int(ln.split()[1]), int(ln.split()[2])) for ln in out if ln.strip()}
    for cname, mirror in pairs:
        last = mirror._fields_[-1][0]
        assert got[cname] == (ctypes.sizeof(mirror), getattr(mirror, last).offset), (cname, got[cname])


def test_product_path_fails_loudly_without_gpu():
    from articulatory_b200 import models as M
    from articulatory_b200._lib import ArticError
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(in_channels=13, channels=32, upsample_scales=[2, 2], upsample_kernel_sizes=[4, 4])
    with pytest.raises(ArticError):
        G(torch.zeros(1, 13, 8))


# --------------------------------------------------------------------------- collater / decode indexing
def test_speech_collater_matches_reference_output():
    """articulatory_b200.data.SpeechCollater == reference SpeechCollater output (tests/golden/collater.npz,
    generated by tests/golden/make_golden.py with np.random.seed(11))."""
    from articulatory_b200.data import SpeechCollater
    gold = np.load(os.path.join(HERE, "golden", "collater.npz"))
    rng = np.random.RandomState(3)
    items = []
    for n_art in (40, 26, 25, 200):
        audio = rng.randn(n_art * 80 + 17).astype(np.float32)
        art = rng.randn(n_art, 13).astype(np.float32)
        items.append({"audio": audio, "art": art})
    gp = {"use_ar": True, "ar_input": 512, "out_channels": 1}
    coll = SpeechCollater(batch_max_steps=2000, hop_size=80, dataset_mode="a2w",
                          config={"generator_params": gp, "batch_max_steps": 2000, "hop_size": 80})
    np.random.seed(11)
    out = coll(items)
    assert np.array_equal(out["x"][0].numpy(), gold["x"])
    assert np.array_equal(out["y"].numpy(), gold["y"])
    assert np.array_equal(out["ar"].numpy(), gold["ar"])
    assert out["x"][0].shape == (3, 13, 25) and out["y"].shape == (3, 1, 2000) and out["ar"].shape == (3, 1, 512)


def test_collater_rejects_off_path_modes():
    from articulatory_b200.data import SpeechCollater
    with pytest.raises(NotImplementedError):
        SpeechCollater(dataset_mode="w2a", config={})
    with pytest.raises(NotImplementedError):
        SpeechCollater(config={"package_mode": "pad"})


def test_chunk_plan_matches_oracle_and_golden():
    from articulatory_b200.decode import chunk_plan
    from oracle import torch_oracle as O
    for n in (1, 24, 25, 26, 57, 99, 100, 101, 637):
        for bms, hop in ((2000, 80), (8000, 80)):
            assert chunk_plan(n, bms, hop) == O.ar_chunk_plan(n, bms, hop)
    assert chunk_plan(57, 2000, 80) == [(0, 25, 0, 2000), (25, 50, 2000, 4000), (50, 57, 4000, 4560)]
    idx = json.load(open(os.path.join(HERE, "golden", "indexing.json")))
    assert idx["mpd_padded"]["8512,3"] == 8514      # fixture sanity (used by the GPU tests)


# --------------------------------------------------------------------------- launcher
def test_launcher_environment_and_commands():
    from articulatory_b200.distributed import launch
    args = launch.parse_args(["--nproc_per_node", "4", "--nnodes", "2", "--node_rank", "1", "--master_port", "12345",
                              "-c", "articulatory-train", "--config", "x.yaml"])
    env = launch.rank_env(args, 2, base={})
    assert env == {"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "12345", "WORLD_SIZE": "8", "RANK": "6",
                   "LOCAL_RANK": "2", "OMP_NUM_THREADS": "1"}
    assert launch.rank_cmd(args, 2) == ["articulatory-train", "--local_rank=2", "--config", "x.yaml"]
    args = launch.parse_args(["--use_env", "-m", "articulatory_b200.bin.train", "--outdir", "o"])
    cmd = launch.rank_cmd(args, 0)
    assert cmd[1:4] == ["-u", "-m", "articulatory_b200.bin.train"] and "--local_rank=0" not in cmd


# --------------------------------------------------------------------------- data parallel on gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    import torch.nn.functional as F

    from articulatory_b200.parallel import DataParallel
    dp = DataParallel(backend="gloo")
    assert (dp.rank, dp.world) == (rank, world)
    torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights ...
    lin = torch.nn.Linear(6, 3)
    dp.broadcast_parameters(lin)                       # ... and end with rank 0's
    torch.manual_seed(7)
    gx, gy = torch.randn(10, 6), torch.randn(10, 3)    # the same global batch on every rank
    shard = dp.shard({"x": gx, "y": gy})
    # per-rank mean loss scaled by 1/world (TrainStep folds 1/world into the loss seeds); ragged shards
    # are weighted by their share so that the sum equals the global-batch mean
    w = shard["x"].shape[0] / gx.shape[0]
    loss = F.mse_loss(lin(shard["x"]), shard["y"]) * w
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in lin.parameters()])
    dp.all_reduce(flat)
    vals = dp.mean_scalars(torch.tensor([float(rank), 1.0]))
    idx = dp.sampler_indices(11, epoch=3)
    torch.save({"w": lin.weight.detach().clone(), "flat": flat, "n": shard["x"].shape[0], "vals": vals, "idx": idx},
               os.path.join(out_dir, f"r{rank}.pt"))
    dp.barrier()
    dp.close()


def test_data_parallel_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    import torch.nn.functional as F
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt", weights_only=False) for i in range(world)]
    assert torch.equal(r[0]["w"], r[1]["w"])                          # broadcast
    assert r[0]["n"] + r[1]["n"] == 10                                # shard covers the batch
    assert torch.equal(r[0]["flat"], r[1]["flat"])                    # all-reduce result identical on all ranks
    torch.manual_seed(100)                                            # rank 0's initial weights
    lin = torch.nn.Linear(6, 3)
    torch.manual_seed(7)
    gx, gy = torch.randn(10, 6), torch.randn(10, 3)
    F.mse_loss(lin(gx), gy).backward()                                # 1-process global-batch gradient
    ref = torch.cat([p.grad.reshape(-1) for p in lin.parameters()])
    assert torch.allclose(r[0]["flat"], ref, rtol=1e-5, atol=1e-7)
    assert torch.allclose(r[0]["vals"], torch.tensor([0.5, 1.0]))
    both = sorted(r[0]["idx"] + r[1]["idx"])
    assert len(r[0]["idx"]) == len(r[1]["idx"]) == 6 and set(both) == set(range(11))


def _dp_wire_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    from articulatory_b200.parallel import DataParallel
    dp = DataParallel(backend="gloo", compress="bf16")
    torch.manual_seed(rank)
    a, b = torch.randn(1000), torch.randn(1000)          # per-rank gradients
    a0, b0 = a.clone(), b.clone()
    dp.all_reduce(a)                                     # nobody asked for the wire buffer: result cast back into `a`
    wire = dp.wire_of(b)                                 # a consumer (FusedAdam.wire) reads the reduced gradient here
    dp.all_reduce(b)
    torch.save({"a0": a0, "b0": b0, "a": a, "b": b, "wire": wire.clone(), "bytes": dp.bytes_reduced,
                "desc": dp.describe()}, os.path.join(out_dir, f"w{rank}.pt"))
    dp.barrier()
    dp.close()


def test_data_parallel_bf16_wire_gloo_world2(tmp_path):
    """DataParallel(compress="bf16") (DDP's bf16_compress_hook semantics): cast, sum in bf16, and either cast back (no
    consumer of the wire buffer) or leave the result in the wire buffer for the optimiser (wire_of) without touching the
    fp32 gradient."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_dp_wire_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"w{i}.pt", weights_only=False) for i in range(world)]
    want_a = (r[0]["a0"].bfloat16() + r[1]["a0"].bfloat16()).float()
    want_b = r[0]["b0"].bfloat16() + r[1]["b0"].bfloat16()
    for i in range(world):
        assert torch.equal(r[i]["a"], want_a)
        assert torch.equal(r[i]["wire"], want_b) and r[i]["wire"].dtype == torch.bfloat16
        assert torch.equal(r[i]["b"], r[i]["b0"])                     # the fp32 buffer keeps the LOCAL gradient
        assert r[i]["bytes"] == 2 * 1000 * 2
        assert r[i]["desc"]["wire_dtype"] == "bf16" and r[i]["desc"]["world"] == 2


def test_mel_filter_ranges_cover_every_nonzero_weight():
    """MelSpectrogramLoss.mel_ranges (the sparse-projection table handed to artic_mel_loss_fwd_bwd): per filter the
    [first, last + 1) bins and per bin the [first, last + 1) filters with a non-zero weight — a weight outside its
    range would silently drop out of the loss."""
    from articulatory_b200.losses import MelSpectrogramLoss
    from oracle import torch_oracle as O
    for params in (O.E2W_MEL_LOSS_PARAMS, dict(fs=22050, fft_size=512, hop_size=128, num_mels=40, fmin=80, fmax=7600)):
        mod = MelSpectrogramLoss(**params)
        mm = mod.melmat.numpy()                                            # (bins, mels)
        rng = mod.mel_ranges.numpy()
        n_bins, n_mels = mm.shape
        assert rng.shape == (2 * n_mels + 2 * n_bins,) and rng.dtype == np.int32
        nnz = 0
        for m in range(n_mels):
            lo, hi = rng[2 * m], rng[2 * m + 1]
            assert 0 <= lo <= hi <= n_bins and not mm[:lo, m].any() and not mm[hi:, m].any()
            nnz += hi - lo
        for k in range(n_bins):
            lo, hi = rng[2 * n_mels + 2 * k], rng[2 * n_mels + 2 * k + 1]
            assert 0 <= lo <= hi <= n_mels and not mm[k, :lo].any() and not mm[k, hi:].any()
        assert nnz < 0.1 * mm.size                                          # triangular filters: a few bins each


def test_batch_prefetcher_order_determinism_errors_and_early_stop():
    """data.BatchPrefetcher: same batches, in order, as the synchronous loop under a seeded np.random (one worker
    draws the collater's windows in job order); worker exceptions surface in the consumer; abandoning the iterator
    stops the worker."""
    import threading

    from articulatory_b200.data import BatchPrefetcher, SpeechCollater, synthetic_utterances
    items = synthetic_utterances(6, n_frames=120, n_feats=5, hop_size=80, seed=3)
    cfg = {"generator_params": {"use_ar": True, "ar_input": 64, "out_channels": 1}}
    col = SpeechCollater(batch_max_steps=1600, hop_size=80, config=cfg)
    groups = [[0, 1], [2, 3], [4, 5], [1, 4]]
    np.random.seed(11)
    want = [col([items[i] for i in g]) for g in groups]
    np.random.seed(11)
    got = list(BatchPrefetcher(lambda g: col([items[i] for i in g]), groups, depth=2, pin=False))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        for k in ("y", "ar"):
            assert torch.equal(a[k], b[k])
        assert torch.equal(a["x"][0], b["x"][0])

    def boom(g):
        if g == 2:
            raise ValueError("bad item")
        return {"y": torch.zeros(1)}
    it = iter(BatchPrefetcher(boom, [0, 1, 2, 3], depth=1, pin=False))
    assert next(it)["y"].shape == (1,) and next(it) is not None
    with pytest.raises(ValueError, match="bad item"):
        next(it)

    n0 = threading.active_count()
    pf = BatchPrefetcher(lambda g: {"y": torch.zeros(1)}, range(1000), depth=1, pin=False)
    for i, _ in enumerate(pf):
        if i == 2:
            break                                   # generator closed -> close() -> worker stops
    pf.close()
    assert not pf._t.is_alive() and threading.active_count() <= n0 + 1


# --------------------------------------------------------------------------- shipped recipe constants
def test_configs_match_shipped_yaml():
    """articulatory_b200.configs (what bench.py / smoke() build the workload from) == the reference's shipped
    egs/ema/voc1/conf/e2w_hifigan.yaml (vendored unchanged under tests/golden/conf/), and == the oracle's constants;
    the package's synthetic batch == the oracle's."""
    import yaml
    from articulatory_b200 import configs as C
    from oracle import torch_oracle as O
    y = yaml.load(open(os.path.join(HERE, "golden", "conf", "e2w_hifigan.yaml")), Loader=yaml.Loader)
    assert y["generator_params"] == C.E2W_GENERATOR_PARAMS == O.E2W_GENERATOR_PARAMS
    assert y["discriminator_params"] == C.E2W_DISCRIMINATOR_PARAMS == O.E2W_DISCRIMINATOR_PARAMS
    assert y["mel_loss_params"] == C.E2W_MEL_LOSS_PARAMS == O.E2W_MEL_LOSS_PARAMS
    assert C.DEFAULT_STFT_LOSS_PARAMS == O.DEFAULT_STFT_LOSS_PARAMS
    cfg = C.e2w_train_config(use_stft_loss=False)
    for k, v in cfg.items():
        if k == "stft_loss_params":          # not in the yaml: MultiResolutionSTFTLoss defaults
            continue
        assert y[k] == v, (k, y[k], v)
    a, b = C.synthetic_batch(3, seed=5), O.synthetic_batch(3, seed=5)
    assert all(torch.equal(a[k], b[k]) for k in ("x", "y", "ar"))
    assert (y["batch_max_steps"], y["hop_size"]) == (8000, 80)
