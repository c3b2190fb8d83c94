#!/usr/bin/env python
"""SM-occupancy timeline of one graph-replayed train step (artic_trace_buffer): every CTA of the
tensor-core conv / weight-gradient kernels records (launch, SM, start, end) with %globaltimer.
Prints: SM-time by kernel kind, the fraction of the step during which each SM holds a tensor-core CTA,
a coarse concurrency histogram, and the longest launches.  Nothing is asserted.
`python tools/sm_timeline.py [--batch 16] [--dump gpurun_out/trace.npy]`"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    import bench
    from articulatory_b200 import _lib
    from articulatory_b200 import models as M
    from articulatory_b200.trainer import TrainStep
    from articulatory_b200 import configs as O

    dev = torch.device("cuda", 0)
    lib = _lib.load()
    cap = 400000
    buf = torch.zeros(1 + 4 * cap, dtype=torch.int64, device=dev)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision=args.precision).to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision=args.precision).to(dev)
    ts = TrainStep(G, D, O.e2w_train_config(use_stft_loss=True), dev)
    b = {k: v.to(dev) for k, v in O.synthetic_batch(args.batch, seed=1234).items()}
    for _ in range(3):
        ts.step(b["x"], b["y"], b["ar"], use_graph=False)
    torch.cuda.synchronize()
    lib.artic_trace_buffer(buf.data_ptr(), cap)          # baked into the graph at capture
    for _ in range(3):
        ts.step(b["x"], b["y"], b["ar"])
    torch.cuda.synchronize()
    buf[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ts.step(b["x"], b["y"], b["ar"])
    e1.record()
    torch.cuda.synchronize()
    lib.artic_trace_buffer(None, 0)
    h = buf.cpu().numpy()
    n = int(min(h[0], cap))
    rec = h[1:1 + 4 * n].reshape(n, 4)
    if args.dump:
        np.save(args.dump, rec)
    launch, kind, sm = rec[:, 0] >> 32, (rec[:, 0] >> 28) & 15, rec[:, 1]
    t0, t1 = rec[:, 2].astype(np.float64), rec[:, 3].astype(np.float64)
    T0, T1 = t0.min(), t1.max()
    span = (T1 - T0) / 1e3
    print(f"step (events) {e0.elapsed_time(e1):.3f} ms; traced span {span / 1e3:.3f} ms; {n} CTAs in "
          f"{len(np.unique(launch))} launches on {len(np.unique(sm))} SMs")
    dur = (t1 - t0) / 1e3
    nsm = int(sm.max()) + 1
    for name, sel in (("conv fwd/dgrad", kind < 8), ("weight gradient", kind == 8)):
        if sel.any():
            print(f"  {name:16s}: {sel.sum():7d} CTAs, SM-time {dur[sel].sum() / 1e3:8.2f} ms = "
                  f"{100 * dur[sel].sum() / (span * nsm):5.1f}% of {nsm} SMs x span; mean CTA {dur[sel].mean():6.1f} us")
    # per-SM union of intervals, and concurrency on a 1 us grid
    grid = np.zeros((nsm, int(span) + 2), dtype=np.int16)
    for s, a, c in zip(sm, ((t0 - T0) / 1e3).astype(int), ((t1 - T0) / 1e3).astype(int)):
        grid[s, a:c + 1] += 1
    busy = (grid > 0).mean(axis=1)
    print(f"  SM holds >= 1 tensor-core CTA: mean {100 * busy.mean():.1f}% of the span (min {100 * busy.min():.1f}, max {100 * busy.max():.1f}); "
          f"two or more co-resident: {100 * (grid > 1).mean():.1f}%")
    act = (grid > 0).sum(axis=0)
    hist, edges = np.histogram(act, bins=[0, 1, 37, 74, 111, 140, 149])
    print("  SMs holding a TC CTA, share of time: " + ", ".join(f"[{edges[i]},{edges[i + 1]}): {100 * hist[i] / act.size:.1f}%" for i in range(len(hist))))
    # per launch: first start, last end, CTA count, mean CTA time
    order = np.argsort(launch, kind="stable")
    ids, starts = np.unique(launch[order], return_index=True)
    rows = []
    for i, lid in enumerate(ids):
        sl = order[starts[i]:starts[i + 1] if i + 1 < len(ids) else None]
        rows.append((lid, int(kind[sl].max()), len(sl), (t0[sl].min() - T0) / 1e3, (t1[sl].max() - T0) / 1e3, dur[sl].mean(),
                     (t0[sl].max() - t0[sl].min()) / 1e3))
    rows.sort(key=lambda r: -(r[4] - r[3]))
    print("  longest launches (id, kind, CTAs, start us, end us, wall us, mean CTA us, start skew us):")
    for r in rows[:25]:
        print(f"    {r[0]:5d} {'wgrad' if r[1] == 8 else 'conv ':5s} {r[2]:4d} {r[3]:9.1f} {r[4]:9.1f} {r[4] - r[3]:7.1f} {r[5]:7.1f} {r[6]:7.1f}")
    walls = np.array([r[4] - r[3] for r in rows])
    ctas = np.array([r[5] for r in rows])
    print(f"  launches: wall mean {walls.mean():.1f} us, median {np.median(walls):.1f}; mean CTA time {ctas.mean():.1f} us; "
          f"sum of launch walls {walls.sum() / 1e3:.2f} ms")


if __name__ == "__main__":
    main()
