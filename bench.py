#!/usr/bin/env python
"""Headline benchmark: audio-samples/sec of one e2w_hifigan G + D + spectral-loss train step
(BASELINE.json metric, configs[1]: full train step on synthetic MNGU0-shape batch 16 per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl ours|reference]

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); the batch is sharded by
utterance (weak scaling: 16 windows per GPU) with an NCCL all-reduce of the flat G and D
gradient buffers.  Prints ONE JSON line on rank 0 (contract in the task description).
`--impl reference` times the reference algorithm on the host cores (the CPU oracle, the only
other place that executes oracle/): /root/reference is pure Python and does not travel to
the GPU box, so the port restated in oracle/torch_oracle.py (pinned to the reference by
tests/golden) stands in for it.
"""
import argparse
import copy
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


def train_config():
    from oracle import torch_oracle as O  # constants only (yaml values restated there)
    return dict(
        use_stft_loss=True, use_mel_loss=True, mel_loss_params=O.E2W_MEL_LOSS_PARAMS,
        stft_loss_params=O.DEFAULT_STFT_LOSS_PARAMS, lambda_aux=45.0, lambda_adv=1.0, lambda_feat_match=2.0,
        use_feat_match_loss=True,
        feat_match_loss_params=dict(average_by_discriminators=False, average_by_layers=False, include_final_outputs=False),
        generator_adv_loss_params=dict(average_by_discriminators=False),
        discriminator_adv_loss_params=dict(average_by_discriminators=False),
        generator_optimizer_params=dict(lr=1e-4, betas=[0.5, 0.9], weight_decay=0.0),
        discriminator_optimizer_params=dict(lr=1e-4, betas=[0.5, 0.9], weight_decay=0.0),
        generator_scheduler_params=dict(gamma=0.5, milestones=[80000, 160000, 240000, 320000]),
        discriminator_scheduler_params=dict(gamma=0.5, milestones=[80000, 160000, 240000, 320000]),
        generator_train_start_steps=1, discriminator_train_start_steps=0,
        generator_grad_norm=-1, discriminator_grad_norm=-1)


def measured_traffic():
    """DRAM bytes per step of the tcgen05 kernels (ncu dram__bytes_{read,write}.sum over one step,
    profiles/r1_traffic.json written by tools/traffic_summary.py); None when not captured."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
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


def cpu_oracle_steps(batch_items, steps, warmup, threads):
    """Reference algorithm (oracle port) on the host: full e2w_hifigan widths, `batch_items` windows."""
    from oracle import torch_oracle as O
    torch.set_num_threads(threads)
    cfg = train_config()
    gsd = O.init_generator_state(O.E2W_GENERATOR_PARAMS, seed=0)
    dsd = O.init_discriminator_state(O.E2W_DISCRIMINATOR_PARAMS, seed=1)
    gopt, dopt = O.AdamState(gsd), O.AdamState(dsd)
    batch = O.synthetic_batch(batch_items)
    times = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            O.train_step(gsd, dsd, O.E2W_GENERATOR_PARAMS, O.E2W_DISCRIMINATOR_PARAMS, gopt, dopt, batch, 2 + it,
                         use_stft_loss=cfg["use_stft_loss"], use_mel_loss=cfg["use_mel_loss"])
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def stft_loss_gpu_ms(dev, reps=50):
    """BASELINE metric, second half: MR-STFT loss forward + backward (three resolutions), B = 16, T = 8000, device
    resident, CUDA events on the launching stream.  Algorithmic bytes = read x, y + write dL/dx = 3 * B * T * 4
    (SURVEY 8d): the figure is launch / latency bound, GB/s is reported against that."""
    from articulatory_b200.losses import MultiResolutionSTFTLoss
    from oracle import torch_oracle as O
    mod = MultiResolutionSTFTLoss(**O.DEFAULT_STFT_LOSS_PARAMS)
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(BATCH_PER_GPU, T, generator=g) * 0.1).to(dev)
    y = (torch.randn(BATCH_PER_GPU, T, generator=g) * 0.1).to(dev)
    R = len(mod.resolutions)
    sums = torch.zeros((R, 3), dtype=torch.float32, device=dev)
    dx = torch.zeros((BATCH_PER_GPU, T), dtype=torch.float32, device=dev)

    def once():
        sums.zero_()
        for r, res in enumerate(mod.resolutions):
            res.forward(x, y, sums[r])
        for r, res in enumerate(mod.resolutions):
            res.backward(x, y, sums[r], 1.0 / R, 1.0 / R, dx)

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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    items = 2
    times = cpu_oracle_steps(items, args.steps, min(args.warmup, 1), threads)
    ms = 1e3 * sum(times) / len(times)
    value = items * T / (ms / 1e3)
    line = {"impl": "reference", "metric": "audio-samples/sec, e2w_hifigan G+D+spectral-loss train step",
            "value": value, "unit": "audio-samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "e2w_hifigan.yaml G+D+mel+MR-STFT train step (bounded sample: 2 windows per step)",
                       "batch_per_step": items, "frames": FRAMES, "samples_per_window": T},
            "cpu_baseline": {"value": value, "unit": "audio-samples/s", "cores": threads, "kind": "port",
                             "sample": f"{items} windows/step x {args.steps} steps, full e2w_hifigan widths, "
                                       "oracle/torch_oracle.py (reference is pure Python, absent on this box)"},
            "e2e": {"value": value, "unit": "audio-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist

    from articulatory_b200 import _lib
    from articulatory_b200 import models as M
    from articulatory_b200.parallel import DataParallel, env_world
    from articulatory_b200.trainer import TrainStep
    from oracle import torch_oracle as O  # synthetic workload generator + yaml constants

    rank, local, world = env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dp = DataParallel(backend="nccl", device=dev)      # one process per GPU; NCCL over NVLink / NVSwitch
    _lib.load()

    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision=args.precision).to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision=args.precision).to(dev)
    dp.broadcast_parameters(G, D)
    # gradient exchange: sum of the flat fp32 gradient buffers (1/world is folded into the loss seeds)
    ts = TrainStep(G, D, train_config(), dev, world_size=world, all_reduce=dp.all_reduce if world > 1 else None)
    B = BATCH_PER_GPU
    host = O.synthetic_batch(B, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    devb = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # schedule gates: steps 0 and 1 are not steady state (bin/train.py:268,350,388)
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
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        ts.step(pinned["x"], pinned["y"], pinned["ar"])
        last = ts.vals.cpu()            # D2H read of the nine logged scalars (forces completion)
    f1.record()
    barrier()
    e2e_ms_total = f0.elapsed_time(f1)
    sampler.stop_flag = True
    t = torch.tensor([ms_total, e2e_ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = t.tolist()
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
    line = {
        "metric": "audio-samples/sec, e2w_hifigan G+D+spectral-loss train step",
        "value": value, "unit": "audio-samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (fp32 storage, split-bf16 tcgen05, fp32 accumulate)", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": "e2w_hifigan.yaml full G+D+mel+MR-STFT train step, synthetic 13-dim 200 Hz EMA -> 16 kHz, "
                               "batch 16 windows of 8000 samples per GPU (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": world * B, "frames": FRAMES, "samples_per_window": T,
                   "parallelism": f"dp{world}", "precision": args.precision,
                   "l2": "per-step activation working set (>2 GB) exceeds the 126 MB L2; no explicit flush",
                   "cuda_graph": ts._graph is not None, "losses": last.tolist() if last is not None else None},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "audio-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 36},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": measured_traffic(),
                     "kernel": "tc::tapconv_tc_kernel + tc::tapwgrad_tc_kernel (tcgen05 implicit-GEMM conv, data and weight gradients)",
                     "note": f"algorithmic 194.70 GFLOP/window (SURVEY 8d: 4 F_G + 8 F_D) x {B} windows / step time, per GPU, "
                             f"timed with CUDA events over the whole step; peak = {peak_src}; traffic = DRAM bytes per step of the "
                             "tensor-core kernels (ncu, profiles/r1_traffic.json)"},
    }
    try:        # second half of BASELINE's metric; never allowed to break the headline line
        sms = stft_loss_gpu_ms(dev)
        nbytes = 3 * B * T * 4
        line["stft_loss"] = {"ms": sms, "shape": f"B={B}, T={T}, 3 resolutions (1024/2048/512), fwd + bwd",
                             "algorithmic_bytes": nbytes, "achieved_GBps": nbytes / (sms * 1e-3) / 1e9,
                             "peak_GBps": peak_gbs, "note": "six launches + one 36-byte memset: latency bound, not HBM bound"}
        if world == 1 and not args.no_cpu_baseline:
            line["stft_loss"]["cpu_ms"] = stft_loss_cpu_ms(os.cpu_count() or 1)
            line["stft_loss"]["cpu_cores"] = os.cpu_count() or 1
    except Exception as ex:  # noqa: BLE001
        line["stft_loss"] = {"error": str(ex)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times = cpu_oracle_steps(2, 2, 1, threads)
        cms = 1e3 * sum(times) / len(times)
        line["cpu_baseline"] = {"value": 2 * T / (cms / 1e3), "unit": "audio-samples/s", "cores": threads, "kind": "port",
                                "sample": "2 windows/step x 2 steps (1 warm-up), full e2w_hifigan widths, fp32, "
                                          "oracle/torch_oracle.py restatement of Trainer._train_step"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("ARTIC_PRECISION", "bf16"), choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
