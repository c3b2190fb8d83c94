#!/usr/bin/env python
"""Where the data-parallel step spends its time: CUDA events at the joints of the overlapped four-graph schedule
(trainer.TrainStep.step), averaged over steady-state steps on every rank.

    torchrun --nproc-per-node 2 tools/dp_timeline.py [--steps 30]

Prints, per rank: generator forward (g1a), the wait for the previous step's tail + the generator phase (g1b), the G
gradient exchange (critical path), the discriminator phase (g2) and, on the exchange stream, the D gradient exchange
and the tail graph (Adam(D) + weight re-materialisation + log assembly), plus how far the tail reaches into the next
step's generator forward."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--fp32-wire", action="store_true")
    args = ap.parse_args()
    import bench
    from articulatory_b200 import configs as C
    from articulatory_b200.parallel import DataParallel, env_world
    rank, local, world = env_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dp = DataParallel(device=dev, compress=None if args.fp32_wire else "bf16")
    ts = bench.build_step("bf16", dev, world, dp)
    b = {k: v.to(dev) for k, v in C.synthetic_batch(bench.BATCH_PER_GPU, seed=1234 + rank).items()}
    bench.timed_steps(ts, b, 5, 3, dp.barrier)
    ts._trace = []
    ms = bench.timed_steps(ts, b, args.steps, 3, dp.barrier)
    torch.cuda.synchronize()
    tr = [t for t in ts._trace if len(t) == 7][4:]
    names = ["g1a generator forward", "wait tail(k-1) + g1b generator phase", "G gradient exchange", "g2 discriminator phase"]
    out = [f"rank {rank}: {ms:.3f} ms/step over {len(tr)} traced steps"]
    for i, n in enumerate(names):
        out.append(f"  {n:40s} {sum(t[i].elapsed_time(t[i + 1]) for t in tr) / len(tr):7.3f} ms")
    out.append(f"  {'[xs] D gradient exchange':40s} {sum(t[4].elapsed_time(t[5]) for t in tr) / len(tr):7.3f} ms")
    out.append(f"  {'[xs] g3 Adam(D) + prep(D) + logs':40s} {sum(t[5].elapsed_time(t[6]) for t in tr) / len(tr):7.3f} ms")
    reach = [a[6].elapsed_time(b_[1]) for a, b_ in zip(tr[:-1], tr[1:])]
    out.append(f"  {'tail end -> end of next g1a (slack)':40s} {sum(reach) / len(reach):7.3f} ms  (negative: g1b waited for the tail)")
    for r in range(world):
        if r == rank:
            print("\n".join(out), flush=True)
        dp.barrier()
    dp.close()


if __name__ == "__main__":
    main()
