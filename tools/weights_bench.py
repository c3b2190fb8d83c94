#!/usr/bin/env python
"""Times (or, under ncu, just runs) the weight-side passes of the discriminator / generator:
prep (weight-norm scale + both prepared layouts), unprep (+ weight-norm backward), Adam.
`python tools/weights_bench.py [--ncu]`"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ncu = "--ncu" in sys.argv
    from articulatory_b200 import models as M
    from articulatory_b200.optim import FusedAdam
    from articulatory_b200 import configs as O
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision="bf16").to(dev)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision="bf16").to(dev)
    for name, mod in (("D", D), ("G", G)):
        opt = FusedAdam(mod, lr=1e-4)
        eng = mod._ensure_ready()
        ws = eng.wset
        n = sum(p.numel() for p in mod.parameters())
        passes = [("prep", ws.prep, 8.0), ("unprep+wn_bwd", lambda: ws.unprep(opt.grad_views), 20.0),
                  ("adam", opt.step, 28.0)]
        for pname, fn, bpp in passes:
            fn()
            torch.cuda.synchronize()
            if ncu:
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 * 1e3
            print(f"{name} {pname:14s} {us:8.1f} us  {n * bpp / us / 1e6:6.2f} TB/s (algorithmic {bpp:.0f} B/param, {n / 1e6:.1f} M params)")


if __name__ == "__main__":
    main()
