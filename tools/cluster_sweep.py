#!/usr/bin/env python
"""Isolated timing of the deep conv layers with / without cluster weight multicast (artic_debug_set key 22:
1 = off, 0 = clusters of 2, 4 = clusters of 4), forward direction, bf16.  Prints one table; nothing is asserted."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import BF16  # noqa: E402
from articulatory_b200.convspec import ConvSpec  # noqa: E402
from articulatory_b200.engine import ConvLayer, SeqT  # noqa: E402

DEV = "cuda:0"
SHAPES = [
    (dict(kind="conv", cin=256, cout=256, k=3, padding=1), 16, 500, 1),
    (dict(kind="conv", cin=256, cout=256, k=7, padding=3), 16, 500, 1),
    (dict(kind="conv", cin=256, cout=256, k=11, padding=5), 16, 500, 1),
    (dict(kind="conv", cin=128, cout=128, k=11, padding=5), 16, 2000, 1),
    (dict(kind="conv", cin=256, cout=512, k=7, padding=3), 16, 100, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=5, padding=2), 64, 53, 2),
    (dict(kind="conv", cin=1024, cout=1024, k=5, padding=2), 352, 10, 11),
    (dict(kind="conv", cin=1024, cout=1024, k=5, padding=2), 32, 34, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=41, padding=20, groups=16), 32, 34, 1),
]


def seq(N, L, C, ni):
    if ni == 1:
        return SeqT((torch.randn(N, L, C, device=DEV) * 0.5).to(torch.bfloat16), N, L, C)
    B = N // ni
    t = (torch.randn(B, L, ni, C, device=DEV) * 0.5).to(torch.bfloat16)
    return SeqT(t, N, L, C, n_inner=ni, s_outer=L * ni * C, s_inner=C, s_row=ni * C)


def timeit(fn, iters=20, reps=3):
    best = 1e9
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3


lib = _lib.load()
print(f"{'layer':58s} {'off us':>8s} {'cs2 us':>8s} {'cs4 us':>8s}  clusters taken (cs2, cs4)")
for kw, N, L, ni in SHAPES:
    spec = ConvSpec(**kw)
    lay = ConvLayer(spec, "l", BF16, BF16)
    lay.bind({"l.weight": torch.randn(spec.weight_shape(), device=DEV) * 0.05, "l.bias": torch.zeros(spec.cout, device=DEV)})
    lay.prep()
    X = seq(N, L, spec.cin, ni)
    Y = X.like(C=spec.cout, L=spec.out_len(L))
    row = []
    taken = []
    for key in (1, 0, 4):
        lib.artic_debug_set(22, key)
        _lib.path_counts(reset=True)
        row.append(timeit(lambda: lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1)))
        taken.append(_lib.path_counts()["conv_tc_cluster"] > 0)
    lib.artic_debug_set(22, 0)
    gf = 2.0 * N * spec.out_len(L) * spec.cout * spec.cig * spec.k / 1e9
    print(f"{str(kw)[:56]:58s} {row[0]:8.1f} {row[1]:8.1f} {row[2]:8.1f}  {taken[1:]}  {gf:.1f} GF  best {gf / min(row) * 1e3:.0f} TF/s")
