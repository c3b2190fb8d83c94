"""Data-parallel train step on two GPUs (NCCL): two ranks with half the batch each must take the same
optimizer steps as one GPU with the whole batch (losses that are plain batch means: mel + adversarial +
feature matching; spectral convergence is a per-rank ratio by design).
Skipped unless two CUDA devices are visible (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`)."""
import os
import socket
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(golden):
    from tests.test_gpu_models import _train_config
    return _train_config(golden, stft=False)


def _build(golden, dev):
    from articulatory_b200 import models as M
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**golden["generator_params"])
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**golden["discriminator_params"])
    G.load_state_dict(golden["gsd"])
    D.load_state_dict(golden["dsd"])
    return G.to(dev), D.to(dev)


def _batch(golden):
    b = golden["batch"]          # 2 items -> 4 items (second pair time-reversed) so each rank gets 2
    return {"x": torch.cat([b["x"], b["x"].flip(2)]), "y": torch.cat([b["y"], b["y"].flip(2)]),
            "ar": torch.cat([b["ar"], b["ar"].flip(2)])}


def _worker(rank, world, port, out_dir, overlap="1", compress=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank), ARTIC_DP_OVERLAP=overlap)
    from articulatory_b200.parallel import DataParallel
    from articulatory_b200.trainer import TrainStep
    golden = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_e2w.pt"),
                        weights_only=False)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dp = DataParallel(backend="nccl", device=dev, compress=compress)
    G, D = _build(golden, dev)
    dp.broadcast_parameters(G, D)
    ts = TrainStep(G, D, _cfg(golden), dev, world_size=world, all_reduce=dp.all_reduce, grad_wire=dp.wire_of)
    assert (ts.optD.wire is not None) == (compress == "bf16")
    shard = {k: v.to(dev) for k, v in dp.shard(_batch(golden)).items()}
    for _ in range(6):
        ts.step(shard["x"], shard["y"], shard["ar"], use_graph=True)
    vals = ts.last_values()          # waits for the overlapped tail (D exchange + Adam(D)) of the last step
    torch.cuda.synchronize()
    assert ts._overlap == (overlap == "1")
    torch.save({"g": {k: v.cpu() for k, v in G.state_dict().items()}, "d": {k: v.cpu() for k, v in D.state_dict().items()},
                "vals": vals}, os.path.join(out_dir, f"r{rank}_{overlap}{compress or ''}.pt"))
    dp.barrier()
    dp.close()


def test_two_gpu_step_matches_single_gpu(golden, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from articulatory_b200.trainer import TrainStep
    from tests.helpers import rel_err
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), "1"), nprocs=2, join=True)
    mp.spawn(_worker, args=(2, port + 1 if port < 65000 else port - 1, str(tmp_path), "0"), nprocs=2, join=True)
    mp.spawn(_worker, args=(2, port + 2 if port < 65000 else port - 2, str(tmp_path), "1", "bf16"), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0_1.pt", weights_only=False)
    r1 = torch.load(tmp_path / "r1_1.pt", weights_only=False)
    s0 = torch.load(tmp_path / "r0_0.pt", weights_only=False)
    for k in r0["g"]:
        assert torch.equal(r0["g"][k], r1["g"][k]), f"ranks diverged on {k}"
    # the overlapped schedule (D exchange + Adam(D) under the next generator forward) is the SAME arithmetic as the
    # blocking one; the split-K weight gradients are summed with atomics, so two runs agree to fp32 summation noise
    # (amplified by Adam's normalisation), not bit for bit
    for name in ("g", "d"):
        for k in r0[name]:
            assert rel_err(r0[name][k], s0[name][k]) < 1e-4, f"overlapped schedule changed {name} parameter {k}"
    for k, v in s0["vals"].items():
        assert abs(r0["vals"][k] - v) <= 1e-4 * abs(v), (k, r0["vals"][k], v)
    # bf16 wire (the bench's exchange in the bf16 mode; Adam consumes the wire buffer directly): the ranks stay in lock
    # step, and six steps land where the exact exchange lands up to the wire's rounding (2^-9 per gradient element)
    w0 = torch.load(tmp_path / "r0_1bf16.pt", weights_only=False)
    w1 = torch.load(tmp_path / "r1_1bf16.pt", weights_only=False)
    for name in ("g", "d"):
        for k in w0[name]:
            assert torch.equal(w0[name][k], w1[name][k]), f"ranks diverged on {k} (bf16 wire)"
            delta = r0[name][k] - golden["gsd" if name == "g" else "dsd"][k]
            if delta.abs().max() > 0:
                assert rel_err(w0[name][k] - golden["gsd" if name == "g" else "dsd"][k], delta) < 0.1, (name, k)
    for k, v in r0["vals"].items():
        assert abs(w0["vals"][k] - v) <= 5e-3 * abs(v), (k, w0["vals"][k], v)
    dev = torch.device("cuda", 0)
    G, D = _build(golden, dev)
    ts = TrainStep(G, D, _cfg(golden), dev)
    full = {k: v.to(dev) for k, v in _batch(golden).items()}
    for _ in range(6):
        ts.step(full["x"], full["y"], full["ar"], use_graph=False)
    for k, v in G.state_dict().items():
        d_ref = v.cpu() - golden["gsd"][k]
        d_dp = r0["g"][k] - golden["gsd"][k]
        if d_ref.abs().max() > 0:
            assert rel_err(d_dp, d_ref) < 0.1, (k, rel_err(d_dp, d_ref))     # Adam's sign-like first steps amplify fp noise
    one = ts.last_values()
    for k in ("train/mel_loss", "train/adversarial_loss", "train/feature_matching_loss"):
        both = 0.5 * (r0["vals"][k] + r1["vals"][k])
        assert abs(both - one[k]) <= 2e-3 * abs(one[k]), (k, both, one[k])
