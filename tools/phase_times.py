#!/usr/bin/env python
"""Device time of each phase of the train step, each captured as its own CUDA graph and
replayed (so the numbers are launch-overhead free and include the intra-phase stream concurrency):

  G fwd (tape) | D fwd [fake|real] 2B | FM/adv seeds + D dgrad (fake half) | spectral losses |
  G bwd | Adam(G)+prep | G fwd (no tape) | D fwd fake B | D bwd 2B (dgrad+wgrad) | Adam(D)+prep

with the algorithmic FLOPs of the phase (F_G = 23.534, F_D = 12.570 GFLOP / item).
`python tools/phase_times.py [--batch 16] [--reps 10]`"""
import argparse
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

F_G, F_D = 23.534, 12.570


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import bench
    from articulatory_b200 import models as M
    from articulatory_b200.engine import fork_join, slice_seq
    from articulatory_b200.trainer import TrainStep
    from articulatory_b200 import configs as O

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision=args.precision).to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision=args.precision).to(dev)
    ts = TrainStep(G, D, O.e2w_train_config(use_stft_loss=True), dev)
    b = {k: v.to(dev) for k, v in O.synthetic_batch(args.batch, seed=1234).items()}
    x, y, ar = b["x"], b["y"], b["ar"]
    B, _, T = y.shape
    for _ in range(3):
        ts.step(x, y, ar, use_graph=False)
    torch.cuda.synchronize()
    engG, engD = G._ensure_ready(), D._ensure_ready()
    st = {}

    def g_fwd():
        st["y_"], st["tapeG"] = engG.forward(x, ar, save=True)

    def d_fwd2():
        st["outs2"], st["tape2"] = engD.forward(None, save=True, parts=(ar, (st["y_"], y)))

    def seeds_dgrad():
        outs2 = st["outs2"]
        outs_f = [[slice_seq(o, 0, B) for o in lst] for lst in outs2]
        douts = [[o.like() for o in lst] for lst in outs_f]
        st["d_in"] = engD.backward(engD.slice_tape(st["tape2"], 0, B), douts, grads=None, need_dx=True)

    def spectral():
        y2d, t2d = st["y_"].reshape(B, T), y.reshape(B, T)
        dy = torch.zeros((B, 1, T), dtype=torch.float32, device=dev)
        st["dy"] = dy
        jobs = []
        for r, res in enumerate(ts.stft.resolutions if ts.use_stft else []):
            def job(r=r, res=res):
                res.forward(y2d, t2d, ts.stft_sums[r])
                res.backward(y2d, t2d, ts.stft_sums[r], 1.0, 1.0, dy)
            jobs.append(job)
        if ts.use_mel:
            def mel_job():
                n = ts.mel.numel(B, T)
                ts.mel.loss_and_grad(y2d, t2d, 1.0 / n, ts.slots[0:], 1.0 / n, dy)
            jobs.append(mel_job)
        fork_join(jobs)

    def g_bwd():
        ts.optG.zero_grad()
        engG.backward(st["tapeG"], st["dy"], ts.optG.grad_views)

    def adam_g():
        ts.optG.step()
        G._ensure_ready()

    def g_fwd_notape():
        st["y2"], _ = engG.forward(x, ar, save=False)

    def d_fwd1():
        engD.forward(None, save=True, into=st["tape2"], lo=0, parts=(ar, (st["y2"],)))

    def d_bwd():
        outs2 = [acts[1:] for acts in st["tape2"]["chains"]]
        ts.optD.zero_grad()
        douts = [[None] * (len(lst) - 1) + [lst[-1].like()] for lst in outs2]
        engD.backward(st["tape2"], douts, grads=ts.optD.grad_views, need_dx=False)

    def adam_d():
        ts.optD.step()
        D._ensure_ready()

    phases = [("G fwd (tape)", g_fwd, F_G * B), ("D fwd [fake|real] 2B", d_fwd2, 2 * F_D * B),
              ("D dgrad fake half", seeds_dgrad, F_D * B), ("spectral losses fwd+bwd", spectral, 0.0),
              ("G bwd (dgrad+wgrad)", g_bwd, 2 * F_G * B), ("Adam(G) + weight prep", adam_g, 0.0),
              ("G fwd (no tape)", g_fwd_notape, F_G * B), ("D fwd fake B", d_fwd1, F_D * B),
              ("D bwd 2B (dgrad+wgrad)", d_bwd, 4 * F_D * B), ("Adam(D) + weight prep", adam_d, 0.0)]
    snap = ts._snapshot()
    total = 0.0
    print(f"{'phase':28s} {'ms':>8s} {'GFLOP':>9s} {'TFLOP/s':>8s}")
    for name, fn, gf in phases:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()                                   # eager once: lazy allocations
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        total += ms
        print(f"{name:28s} {ms:8.3f} {gf:9.1f} {gf / ms if gf else 0.0:8.1f}")
        ts._restore(snap)
        engG, engD = G._ensure_ready(), D._ensure_ready()
    print(f"{'sum':28s} {total:8.3f}")


if __name__ == "__main__":
    main()
