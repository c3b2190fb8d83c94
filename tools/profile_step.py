#!/usr/bin/env python
"""Run the bench workload eagerly (no CUDA graph) and bracket ONE steady-state train step with
cudaProfilerStart/Stop, for `ncu --profile-from-start off ...` launch lists and captures:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--precision bf16]
"""
import argparse
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
    args = ap.parse_args()
    from articulatory_b200 import configs as O
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
    torch.cuda.cudart().cudaProfilerStart()
    ts.step(b["x"], b["y"], b["ar"], use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("losses", ts.last_values())


if __name__ == "__main__":
    main()
