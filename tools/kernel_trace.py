#!/usr/bin/env python
"""Full kernel timeline (CUPTI through torch.profiler) of one graph-replayed train step: writes
(name, start us, duration us, stream) of every kernel to a compact JSON for offline analysis.
`python tools/kernel_trace.py --out gpurun_out/ktrace.json`"""
import argparse
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default="gpurun_out/ktrace.json")
    args = ap.parse_args()
    import bench
    from articulatory_b200 import models as M
    from articulatory_b200.trainer import TrainStep
    from articulatory_b200 import configs as O
    from torch.profiler import ProfilerActivity, profile

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision=args.precision).to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision=args.precision).to(dev)
    ts = TrainStep(G, D, O.e2w_train_config(use_stft_loss=True), dev)
    b = {k: v.to(dev) for k, v in O.synthetic_batch(args.batch, seed=1234).items()}
    for _ in range(6):
        ts.step(b["x"], b["y"], b["ar"])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ts.step(b["x"], b["y"], b["ar"])
        torch.cuda.synchronize()
    tmp = args.out + ".chrome.json"
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    ks = [(e["name"], e["ts"], e["dur"], e.get("args", {}).get("stream", -1)) for e in ev
          if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ks.sort(key=lambda k: k[1])
    t0 = ks[0][1] if ks else 0
    out = [[n[:80], round(t - t0, 2), round(d, 2), s] for n, t, d, s in ks]
    json.dump(out, open(args.out, "w"))
    os.remove(tmp)
    span = (max(t + d for _, t, d, _ in out) if out else 0)
    print(f"{len(out)} kernels, span {span / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
