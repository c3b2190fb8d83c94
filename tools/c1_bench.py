#!/usr/bin/env python
"""Times the channel-1 kernels on the first-layer shapes of the discriminators (MSD scale 1: 1 -> 128, k = 15 on
a (32, 8512) batch): forward (Cin = 1), data gradient (Cout = 1) and weight gradient.  Nothing is asserted."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import BF16, F32  # noqa: E402
from articulatory_b200.convspec import ConvSpec  # noqa: E402
from articulatory_b200.engine import ConvLayer, SeqT  # noqa: E402
from tools.tc_sweep import timeit  # noqa: E402

DEV = "cuda:0"


def main():
    _lib.load()
    for (N, L, cout, k) in ((32, 8512, 128, 15), (16, 8512, 128, 15), (32, 4256, 128, 15)):
        spec = ConvSpec(kind="conv", cin=1, cout=cout, k=k, padding=(k - 1) // 2)
        lay = ConvLayer(spec, "l", F32, BF16)
        w = torch.randn(spec.weight_shape(), device=DEV) * 0.05
        lay.bind({"l.weight": w, "l.bias": torch.zeros(cout, device=DEV)})
        lay.prep()
        X = SeqT(torch.randn(N, L, 1, device=DEV), N, L, 1)
        Y = SeqT((torch.randn(N, L, cout, device=DEV) * 0.5).to(torch.bfloat16), N, L, cout)
        dX = SeqT(torch.zeros(N, L, 1, device=DEV), N, L, 1)
        grads = {"l.weight": torch.zeros_like(w), "l.bias": torch.zeros(cout, device=DEV)}
        t_f = timeit(lambda: lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1))
        t_d = timeit(lambda: lay.dgrad(Y, dX=dX))
        t_w = timeit(lambda: lay.wgrad(X, Y, grads))
        mb = N * L * cout * 2 / 1e6
        print(f"N={N} L={L} 1->{cout} k={k}: wide tensor {mb:.0f} MB | fwd {t_f:6.1f} us ({mb / t_f:.2f} TB/s) | "
              f"dgrad {t_d:6.1f} us ({mb / t_d:.2f} TB/s) | wgrad+bias {t_w:6.1f} us ({mb / t_w:.2f} TB/s)")


if __name__ == "__main__":
    main()
