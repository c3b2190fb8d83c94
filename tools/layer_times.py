#!/usr/bin/env python
"""Per-layer device times of one steady-state train step (eager, CUDA events around every
ConvLayer.forward / dgrad / wgrad call), with the layer's algorithmic FLOPs and the achieved
TFLOP/s.  `python tools/layer_times.py [--precision bf16] [--top 40]`"""
import argparse
import collections
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
    ap.add_argument("--top", type=int, default=45)
    args = ap.parse_args()
    from articulatory_b200 import configs as O
    from articulatory_b200 import engine
    from articulatory_b200 import models as M
    from articulatory_b200.trainer import TrainStep

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision=args.precision).to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision=args.precision).to(dev)
    ts = TrainStep(G, D, O.e2w_train_config(use_stft_loss=True), dev)
    b = {k: v.to(dev) for k, v in O.synthetic_batch(args.batch, seed=1234).items()}
    for _ in range(4):
        ts.step(b["x"], b["y"], b["ar"], use_graph=False)
    torch.cuda.synchronize()

    records = []
    CL = engine.ConvLayer

    def wrap(name, flops_of):
        orig = getattr(CL, name)

        def f(self, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(self, *a, **k)
            e1.record()
            records.append((self, name, flops_of(self, a, k), e0, e1))
            return r
        setattr(CL, name, f)

    def fl_fwd(self, a, k):
        X = a[0]
        s = self.spec
        lout = s.out_len(X.L)
        rows = lout if s.kind != "convT" else X.L
        return 2.0 * X.N * rows * s.cog * s.cig * s.groups * s.k, (X.N, X.L, s.cin, s.cout, s.k, s.stride, s.dilation, s.groups)

    def fl_dgrad(self, a, k):
        dY = a[0]
        s = self.spec
        rows = dY.L if s.kind != "convT" else (k.get("dX") or k.get("dX2")).L
        return 2.0 * dY.N * rows * s.cog * s.cig * s.groups * s.k, (dY.N, dY.L, s.cin, s.cout, s.k, s.stride, s.dilation, s.groups)

    def fl_wgrad(self, a, k):
        X, dY = a[0], a[1]
        s = self.spec
        rows = dY.L if s.kind != "convT" else X.L
        return 2.0 * X.N * rows * s.cog * s.cig * s.groups * s.k, (X.N, X.L, s.cin, s.cout, s.k, s.stride, s.dilation, s.groups)

    wrap("forward", fl_fwd)
    wrap("dgrad", fl_dgrad)
    wrap("wgrad", fl_wgrad)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    ts.step(b["x"], b["y"], b["ar"], use_graph=False)
    t1.record()
    torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
    tot_ms = tot_fl = 0.0
    for lay, name, (fl, shape), e0, e1 in records:
        ms = e0.elapsed_time(e1)
        key = (name, shape)
        agg[key][0] += ms
        agg[key][1] += fl
        agg[key][2] += 1
        tot_ms += ms
        tot_fl += fl
    print(f"eager step {t0.elapsed_time(t1):.2f} ms; conv calls {len(records)}: {tot_ms:.2f} ms, {tot_fl / 1e12:.3f} TFLOP "
          f"-> {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")
    print("   ms     n   TFLOP/s  GFLOP  dir      (N, L, cin, cout, k, stride, dil, groups)")
    for key, (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[: args.top]:
        print(f"{ms:7.3f} {n:4d} {fl / ms / 1e9:8.1f} {fl / 1e9:7.1f}  {key[0]:8s} {key[1]}")


if __name__ == "__main__":
    main()
