#!/usr/bin/env python
"""Headline benchmark: audio-samples/sec of one e2w_hifigan G + D + spectral-loss train step
(BASELINE.json metric, configs[1]: full train step on a synthetic MNGU0-shape batch of 16 windows per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|bf16x3|fp32] [--impl ours|reference]

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); the batch is sharded by utterance (weak
scaling: 16 windows per GPU) with an NCCL all-reduce of the flat G and D gradient buffers.  Prints ONE JSON line
on rank 0 (contract in the task description).

Precision modes.  The HEADLINE (`value`, `e2e`, `roofline`) is `--precision bf16` — BASELINE.json configs[1] names
a bf16 train step: bf16 storage, tcgen05 contraction, fp32 accumulate.  Its error against the fp32 reference is
~6e-3 on the waveform and <= 2e-3 on the nine losses (tests/test_gpu_models.py::test_full_width_*, recorded).  The
PARITY-GATED tensor-core mode is `bf16x3` (fp32 storage, split-bf16 tcgen05; <= 1e-4 on four full-width train
steps against the oracle): the same line reports its step time under `parity_mode`, and `mode_agreement` holds the
relative difference of the nine logged losses between the two modes after identical steps.

Baselines in the line.  `cpu_baseline` / `--impl reference`: the reference's algorithm on the host cores
(oracle/torch_oracle.py, pinned to the reference by tests/golden; /root/reference is pure Python and does not
travel to the GPU box), SAME workload (16 windows per step).  `stock_torch_gpu`: the same oracle modules executed
by stock PyTorch on this B200 (cuDNN / cuBLAS / cuFFT; fp32 and bf16 autocast) — the kernel-quality bar.
The timed GPU arm imports nothing from oracle/.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH_PER_GPU = 16
FRAMES, HOP, AR_LEN = 100, 80, 512
T = FRAMES * HOP
# algorithmic conv/linear FLOPs of one train step per batch item
# = 4*F_G + 8*F_D = 4*23.534 + 8*12.570 GFLOP
FLOP_PER_ITEM = 194.70e9
METRIC = "audio-samples/sec, e2w_hifigan G+D+spectral-loss train step"


def workload_config(world):
    """The `config` object: identical for `--impl ours` and `--impl reference`."""
    return {"workload": "e2w_hifigan.yaml full G+D+mel+MR-STFT train step, synthetic 13-dim 200 Hz EMA -> 16 kHz, "
                        "batch 16 windows of 8000 samples per GPU (BASELINE configs[1])",
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": world * BATCH_PER_GPU, "frames": FRAMES,
            "samples_per_window": T, "parallelism": f"dp{world}"}


def measured_traffic():
    """DRAM bytes per step of the tcgen05 kernels from the ncu capture of THIS round's code
    (profiles/r2_traffic.json, written by tools/traffic_summary.py); None when not captured."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(path))["tensor_core_kernels"]["dram_bytes_per_step"]
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# reference algorithm (oracle) legs: CPU baseline, --impl reference, stock torch on the GPU
# ------------------------------------------------------------------------------------------------------------
def oracle_steps(batch_items, steps, warmup, device="cpu", autocast=False):
    """`steps` timed oracle train steps (reference Trainer._train_step restated) at full e2w_hifigan width."""
    from oracle import torch_oracle as O
    gsd = {k: v.to(device) for k, v in O.init_generator_state(O.E2W_GENERATOR_PARAMS, seed=0).items()}
    dsd = {k: v.to(device) for k, v in O.init_discriminator_state(O.E2W_DISCRIMINATOR_PARAMS, seed=1).items()}
    gopt, dopt = O.AdamState(gsd), O.AdamState(dsd)
    batch = {k: v.to(device) for k, v in O.synthetic_batch(batch_items).items()}
    cuda = str(device).startswith("cuda")
    times = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(warmup + steps):
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and cuda):
                O.train_step(gsd, dsd, O.E2W_GENERATOR_PARAMS, O.E2W_DISCRIMINATOR_PARAMS, gopt, dopt, batch, 2 + it,
                             use_stft_loss=True, use_mel_loss=True)
            if cuda:
                torch.cuda.synchronize()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def stock_torch_gpu(dev):
    """The oracle's modules executed by stock PyTorch on this GPU (cuDNN benchmark on, as reference bin/train.py:1451)."""
    out = {"note": "oracle/torch_oracle.py train_step (the reference's torch ops) on the same B200, B = 16: cuDNN / cuBLAS / "
                   "cuFFT, per-step host syncs as in the reference; 3 warm-up + 5 timed steps"}
    torch.backends.cudnn.benchmark = True
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        try:
            ts = oracle_steps(BATCH_PER_GPU, 5, 3, device=dev, autocast=ac)
            ms = 1e3 * sum(ts) / len(ts)
            out[name] = {"ms_per_step": ms, "value": BATCH_PER_GPU * T / (ms / 1e3), "unit": "audio-samples/s"}
        except Exception as ex:  # noqa: BLE001
            out[name] = {"error": str(ex)[:300]}
        torch.cuda.empty_cache()
    return out


def stft_loss_gpu_ms(dev, reps=50):
    """BASELINE metric, second half: MR-STFT loss forward + backward (three resolutions), B = 16, T = 8000, device
    resident, CUDA events on the launching stream.  Algorithmic bytes = read x, y + write dL/dx = 3 * B * T * 4
    (SURVEY 8d): the figure is launch / latency bound, GB/s is reported against that."""
    from articulatory_b200.configs import DEFAULT_STFT_LOSS_PARAMS
    from articulatory_b200.losses import MultiResolutionSTFTLoss
    mod = MultiResolutionSTFTLoss(**DEFAULT_STFT_LOSS_PARAMS)
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(BATCH_PER_GPU, T, generator=g) * 0.1).to(dev)
    y = (torch.randn(BATCH_PER_GPU, T, generator=g) * 0.1).to(dev)
    R = len(mod.resolutions)
    sums = torch.zeros((R, 3), dtype=torch.float32, device=dev)
    dx = torch.zeros((BATCH_PER_GPU, T), dtype=torch.float32, device=dev)

    from articulatory_b200.engine import fork_join

    def job(r, res):
        res.forward(x, y, sums[r])
        res.backward(x, y, sums[r], 1.0 / R, 1.0 / R, dx)

    def once():         # as in the train step: the three resolutions on three streams, forward then backward each
        sums.zero_()
        fork_join([lambda r=r, res=res: job(r, res) for r, res in enumerate(mod.resolutions)])

    for _ in range(5):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def stft_loss_cpu_ms(threads, reps=3):
    """The same loss through the reference's arithmetic (oracle, torch CPU autograd) on the host cores."""
    from oracle import torch_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(BATCH_PER_GPU, T, generator=g) * 0.1).requires_grad_(True)
    y = torch.randn(BATCH_PER_GPU, T, generator=g) * 0.1
    times = []
    for it in range(reps + 1):
        t0 = time.perf_counter()
        sc, mag = O.mr_stft_loss(x, y)
        (sc + mag).backward()
        x.grad = None
        if it > 0:
            times.append(time.perf_counter() - t0)
    return 1e3 * sum(times) / len(times)


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port) on all host cores, SAME workload as the GPU arm:
    16 windows per step.  Under torchrun only rank 0 runs; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    items = BATCH_PER_GPU
    times = oracle_steps(items, args.steps, max(min(args.warmup, 1), 1))
    ms = 1e3 * sum(times) / len(times)
    value = items * T / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "audio-samples/s", "cores": threads, "kind": "port",
                             "sample": f"{items} windows/step x {args.steps} steps (1 warm-up), full e2w_hifigan widths, fp32, "
                                       "oracle/torch_oracle.py (reference is pure Python, absent on this box); one rank's "
                                       "batch — the value is per-sample normalised"},
            "e2e": {"value": value, "unit": "audio-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def build_step(precision, dev, world, dp, seed=0):
    from articulatory_b200 import configs as C
    from articulatory_b200 import models as M
    from articulatory_b200.trainer import TrainStep
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**C.E2W_GENERATOR_PARAMS, precision=precision).to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**C.E2W_DISCRIMINATOR_PARAMS, precision=precision).to(dev)
    if dp is not None:
        dp.broadcast_parameters(G, D)
    # gradient exchange: sum of the flat gradient buffers (1/world is folded into the loss seeds)
    return TrainStep(G, D, C.e2w_train_config(use_stft_loss=True), dev, world_size=world,
                     all_reduce=dp.all_reduce if (dp is not None and world > 1) else None,
                     grad_wire=dp.wire_of if (dp is not None and world > 1) else None)


def timed_steps(ts, b, steps, warmup, barrier):
    """`steps` graph-replayed steady-state steps on device-resident inputs, CUDA events, ms per step (this rank)."""
    # schedule gates: steps 0 and 1 are not steady state (bin/train.py:268,350,388)
    while ts.steps < 2:
        ts.step(b["x"], b["y"], b["ar"], use_graph=False)
    for _ in range(max(warmup, 3) + 1):            # the first one captures the graphs
        ts.step(b["x"], b["y"], b["ar"])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ts.step(b["x"], b["y"], b["ar"])
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def mode_agreement(dev, n_steps=4):
    """Nine logged losses of bf16 vs bf16x3 after identical steps from identical weights (B = 16, eager): the speed
    mode's error measured against the parity-gated mode (which the tests pin to the oracle at 1e-3)."""
    from articulatory_b200 import configs as C
    from articulatory_b200.trainer import LOG_KEYS
    b = {k: v.to(dev) for k, v in C.synthetic_batch(BATCH_PER_GPU, seed=4321).items()}
    vals = {}
    for prec in ("bf16x3", "bf16"):
        ts = build_step(prec, dev, 1, None, seed=3)
        per = []
        for _ in range(n_steps):
            ts.step(b["x"], b["y"], b["ar"], use_graph=False)
            per.append(ts.values_tensor().cpu().tolist())
        vals[prec] = per
        del ts
        torch.cuda.empty_cache()
    rel = [[abs(a - r) / max(abs(r), 1e-12) for a, r in zip(sa, sr)] for sa, sr in zip(vals["bf16"], vals["bf16x3"])]
    worst = {k: max(step[i] for step in rel[2:]) for i, k in enumerate(LOG_KEYS)}
    return {"steps": n_steps, "worst_rel_diff_bf16_vs_bf16x3": max(worst.values()), "per_loss": worst,
            "note": "bf16x3 is gated at 1e-3 against the oracle (tests/test_gpu_models.py::test_full_width_train_steps_vs_oracle)"}


def car_inference(dev, precision, iters=3):
    """BASELINE configs[2]: e2w_hifigan_car.yaml chunked-AR inference, 32 utterances x 600 frames (24 lock-step
    chunks of 25 frames), pinned host features in, host waveforms out."""
    from articulatory_b200 import configs as C
    from articulatory_b200 import models as M
    from articulatory_b200.decode import BatchedARDecoder
    gp = dict(C.E2W_GENERATOR_PARAMS, final_scale=80, extra_art=False)      # e2w_hifigan_car.yaml:35-58
    cfg = {"generator_params": gp, "batch_max_steps": 2000, "hop_size": 80, "sampling_rate": 16000}
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**gp, precision=precision)
    G.remove_weight_norm()
    G = G.eval().to(dev)
    g = torch.Generator().manual_seed(1234)
    batch, frames = 32, 600
    pinned = [torch.randn(frames, 13, generator=g).pin_memory() for _ in range(batch)]
    dec = BatchedARDecoder(G, cfg)
    dec.decode(pinned)                                  # captures the chunk graph
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        outs = dec.decode(pinned)
        host = [o.cpu() for o in outs]                  # noqa: F841  D2H of the waveforms (end to end)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    n = batch * frames * HOP
    n_chunks = -(-frames // 25)
    flop = 5.884e9 * batch * n_chunks                   # F_G of a 25-frame window
    peak_tf, _, _ = peaks()
    return {"metric": "audio-samples/sec, e2w_hifigan_car chunked-AR inference (BASELINE configs[2])",
            "value": n / (ms / 1e3), "unit": "audio-samples/s", "ms_per_batch": ms, "ms_per_chunk": ms / n_chunks,
            "rtf": (ms / 1e3) / (n / 16000.0), "precision": precision,
            "workload": f"{batch} utterances x {frames} frames, {n_chunks} sequential chunks of 25 frames, CUDA graph per chunk, "
                        "pinned host features in, host waveforms out",
            "roofline": {"bound": "tensor", "achieved": flop / (ms / 1e3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": flop / (ms / 1e3) / 1e12 / peak_tf,
                         "note": "5.884 GFLOP per 25-frame window (SURVEY 8d) x 32 x 24; the chunk is a chain of ~80 small "
                                 "dependent launches: latency bound"}}


def inversion(dev, cpu=True, iters=5):
    """BASELINE configs[4]: speech-to-EMA inversion encoder forward (BiGRU 2 x 256 bidirectional -> 12-dim EMA),
    batch 128 on one B200.  Input = HuBERT-large-sized features (1024-d at 200 Hz, reference
    egs/ema/voc1/local/predict_ema.py:83-90), 2-second utterances (400 frames); pinned host features in, host EMA out."""
    from articulatory_b200.models import BiGRU
    N, C, Tn, H = 128, 1024, 400, 256
    torch.manual_seed(0)
    m = BiGRU(in_channels=C, hidden_size=H, out_channels=12)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.eval().to(dev)
    g = torch.Generator().manual_seed(5)
    host = torch.randn(N, C, Tn, generator=g).pin_memory()
    for _ in range(2):
        m(host.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = m(host.to(dev, non_blocking=True)).cpu()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xd = host.to(dev)
    d0.record()
    for _ in range(iters):
        m(xd)
    d1.record()
    torch.cuda.synchronize()
    ms_dev = d0.elapsed_time(d1) / iters
    frames = N * Tn
    flop = frames * (2 * C * 6 * H + 2 * (2 * H) * 6 * H + 2 * 2 * (2 * H * 3 * H) + 2 * 2 * H * 12)
    peak_tf, _, _ = peaks()
    out = {"metric": "EMA frames/sec, speech-to-EMA inversion encoder forward (BASELINE configs[4])",
           "value": frames / (ms_dev / 1e3), "unit": "frames/s", "ms_per_batch": ms_dev,
           "e2e": {"value": frames / (ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": host.numel() * 4,
                   "d2h_bytes_per_step": y.numel() * 4},
           "audio_seconds_per_second": frames / 200.0 / (ms_dev / 1e3),
           "workload": f"{N} utterances x {Tn} frames (2 s at 200 Hz) of {C}-d features -> 12-dim EMA, hidden {H}, bf16x3 input "
                       "projections on tcgen05 + persistent cluster GRU kernel (fp32 FFMA, W_hh resident in shared memory)",
           "roofline": {"bound": "tensor", "achieved": flop / (ms_dev / 1e3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": flop / (ms_dev / 1e3) / 1e12 / peak_tf,
                        "note": "6.3 MFLOP per frame; 2 x 400 strictly sequential recurrent steps bound the latency, and the "
                                "recurrence runs in fp32 FFMA for parity (the fp32 CUDA-core peak, not the tensor peak, bounds it)"}}
    if cpu:
        from oracle import inversion_oracle as I
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        xs = host[:8].clone()
        I.bigru_forward(sd, xs[:2])
        t0 = time.perf_counter()
        ref = I.bigru_forward(sd, xs)
        dt = time.perf_counter() - t0
        err = float((y[:8] - ref).norm() / ref.norm())
        out["cpu_baseline"] = {"value": 8 * Tn / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": "8 utterances x 400 frames, oracle/inversion_oracle.py (fp32)"}
        out["rel_err_vs_oracle"] = err
    return out


def run_ours(args):
    import torch.distributed as dist

    from articulatory_b200 import _lib
    from articulatory_b200 import configs as C
    from articulatory_b200.parallel import DataParallel, env_world

    rank, local, world = env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # (NCCL_DEBUG=INFO with NCCL_DEBUG_FILE set by the caller lets DataParallel.describe() name NVLS vs ring; it is not
    # set here because NCCL then also prints its version banner on stdout, next to the one JSON line)
    # one process per GPU; NCCL over NVLink / NVSwitch; the bf16 speed mode ships its gradients as bf16
    dp = DataParallel(backend="nccl", device=dev, compress="bf16" if args.precision == "bf16" and not args.fp32_wire else None)
    _lib.load()

    B = BATCH_PER_GPU
    host = C.synthetic_batch(B, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    devb = {k: v.to(dev) for k, v in host.items()}
    ts = build_step(args.precision, dev, world, dp)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ts.step(devb["x"], devb["y"], devb["ar"], use_graph=False)
    ts.step(devb["x"], devb["y"], devb["ar"], use_graph=False)
    c0 = _lib.launch_count
    ts.step(devb["x"], devb["y"], devb["ar"], use_graph=True)        # captures the graphs
    launches_per_step = (_lib.launch_count - c0) // 2 if ts._graph is not None else 0   # eager warm-up + capture
    for _ in range(max(args.warmup, 3)):
        ts.step(devb["x"], devb["y"], devb["ar"])

    # ---- device-resident timing ------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ts.step(devb["x"], devb["y"], devb["ar"])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    # ---- end-to-end: pinned host inputs in, per-step loss read-back out ----------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        ts.step(pinned["x"], pinned["y"], pinned["ar"])
        last = ts.values_tensor().cpu()            # D2H read of the nine logged scalars (forces completion)
    f1.record()
    barrier()
    e2e_ms_total = f0.elapsed_time(f1)
    sampler.stop_flag = True
    t = torch.tensor([ms_total, e2e_ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = t.tolist()
    comm = dp.describe() if hasattr(dp, "describe") else None
    del ts
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3] point: 8 windows per GPU (global 64 on 8 GPUs), every rank takes part --------
    small = None
    if not args.no_extras:
        try:
            db = {k: v.to(dev) for k, v in C.synthetic_batch(8, seed=99 + rank).items()}
            ts8 = build_step(args.precision, dev, world, dp)
            ms8 = timed_steps(ts8, db, max(args.steps // 2, 5), 3, barrier)
            t8 = torch.tensor([ms8], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t8, op=dist.ReduceOp.MAX)
            ms8 = t8.item()
            small = {"workload": "same step, 8 windows per GPU (BASELINE configs[3]: global batch 64 on 8 GPUs)",
                     "batch_per_gpu": 8, "global_batch": 8 * world, "ms_per_step": ms8,
                     "value": world * 8 * T / (ms8 / 1e3), "unit": "audio-samples/s"}
            del ts8
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            small = {"error": str(ex)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms = ms_total / args.steps
    value = world * B * T / (ms / 1e3)
    e2e_value = world * B * T / (e2e_ms_total / args.steps / 1e3)
    peak_tf, peak_gbs, peak_src = peaks()
    achieved_tf = FLOP_PER_ITEM * B / (ms / 1e3) / 1e12          # per GPU
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    dtype_names = {"bf16": "bf16", "bf16x3": "bf16x3 (fp32 storage, split-bf16 tcgen05, fp32 accumulate)", "fp32": "f32"}
    line = {
        "metric": METRIC, "value": value, "unit": "audio-samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_names[args.precision], "data": "synthetic", "config": workload_config(world),
        "run": {"precision": args.precision, "cuda_graph": True,
                "l2": "per-step activation working set (>2 GB) exceeds the 126 MB L2; no explicit flush",
                "losses": last.tolist() if last is not None else None, "comm": comm},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "audio-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 36},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": measured_traffic(),
                     "kernel": "tc::tapconv_tc_kernel + tc::tapwgrad_tc_kernel (tcgen05 implicit-GEMM conv, data and weight gradients)",
                     "note": f"algorithmic 194.70 GFLOP/window (SURVEY 8d: 4 F_G + 8 F_D) x {B} windows / step time, per GPU, "
                             f"timed with CUDA events over the whole step; peak = {peak_src}; traffic = DRAM bytes per step of the "
                             "tensor-core kernels (ncu capture of this round's code, profiles/r2_traffic.json) or null"},
    }
    if small is not None:
        line["batch8_per_gpu"] = small
    if args.no_extras or world > 1:        # the remaining extras are single-GPU measurements
        print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    def guarded(key, fn):
        try:        # extras are never allowed to break the headline line
            line[key] = fn()
        except Exception as ex:  # noqa: BLE001
            line[key] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
        torch.cuda.empty_cache()

    if args.precision == "bf16":
        def parity_mode():
            ts3 = build_step("bf16x3", dev, 1, None)
            ms3 = timed_steps(ts3, devb, max(args.steps // 2, 5), 3, barrier)
            tf3 = FLOP_PER_ITEM * B / (ms3 / 1e3) / 1e12
            return {"precision": "bf16x3", "ms_per_step": ms3, "value": B * T / (ms3 / 1e3), "unit": "audio-samples/s",
                    "roofline_frac": tf3 / peak_tf,
                    "note": "parity-gated tensor-core mode: fp32 storage, x_hi*w_hi + x_hi*w_lo + x_lo*w_hi on tcgen05 (3x the "
                            "MMAs, 2-4x the bytes); <= 1e-4 on the nine losses over four full-width steps vs the oracle"}
        guarded("parity_mode", parity_mode)
        guarded("mode_agreement", lambda: mode_agreement(dev))

    def stft():
        sms = stft_loss_gpu_ms(dev)
        nbytes = 3 * B * T * 4
        d = {"ms": sms, "shape": f"B={B}, T={T}, 3 resolutions (1024/2048/512), fwd + bwd",
             "algorithmic_bytes": nbytes, "achieved_GBps": nbytes / (sms * 1e-3) / 1e9, "peak_GBps": peak_gbs,
             "note": "six launches on three streams (one per resolution, as in the train step) + one 36-byte memset: latency bound, not HBM bound"}
        if not args.no_cpu_baseline:
            d["cpu_ms"] = stft_loss_cpu_ms(os.cpu_count() or 1)
            d["cpu_cores"] = os.cpu_count() or 1
        return d
    guarded("stft_loss", stft)
    guarded("car_inference", lambda: car_inference(dev, args.precision))
    guarded("inversion", lambda: inversion(dev, cpu=not args.no_cpu_baseline))
    if not args.no_cpu_baseline:
        guarded("stock_torch_gpu", lambda: stock_torch_gpu(dev))

        def cpu():
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            times = oracle_steps(BATCH_PER_GPU, 2, 1)
            cms = 1e3 * sum(times) / len(times)
            return {"value": BATCH_PER_GPU * T / (cms / 1e3), "unit": "audio-samples/s", "cores": threads, "kind": "port",
                    "sample": f"{BATCH_PER_GPU} windows/step x 2 steps (1 warm-up), full e2w_hifigan widths, fp32, "
                              "oracle/torch_oracle.py restatement of Trainer._train_step"}
        guarded("cpu_baseline", cpu)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("ARTIC_PRECISION", "bf16"), choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the legs that execute oracle/ (CPU + stock torch GPU)")
    ap.add_argument("--no-extras", action="store_true", help="headline line only")
    ap.add_argument("--fp32-wire", action="store_true", help="exchange gradients in fp32 also in the bf16 mode")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
