#!/usr/bin/env python
"""Tile-shape sweep of the tcgen05 tap-gather kernels on the layer shapes of the e2w_hifigan train
step: for every shape, times forward / dgrad for each forced (bn, mt) (artic_debug_set keys 2, 3)
next to the planner's own choice, and the weight gradient for each forced bn (key 6).
Prints one table; nothing is asserted.  `timeout 600 python tools/tc_sweep.py [fwd|dgrad|wgrad ...]`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import BF16  # noqa: E402
from articulatory_b200.convspec import ConvSpec  # noqa: E402
from articulatory_b200.engine import ConvLayer, SeqT  # noqa: E402

DEV = "cuda:0"
# (spec, N sequences, L_in, n_inner)
SHAPES = []
for C, L in ((256, 500), (128, 2000), (64, 4000), (32, 8000)):
    for k in (3, 7, 11):
        SHAPES.append((dict(kind="conv", cin=C, cout=C, k=k, padding=(k - 1) // 2), 16, L, 1))
    SHAPES.append((dict(kind="conv", cin=C, cout=C, k=11, dilation=5, padding=25), 16, L, 1))
SHAPES += [
    (dict(kind="convT", cin=512, cout=256, k=10, stride=5, padding=3, output_padding=1), 16, 100, 1),
    (dict(kind="convT", cin=256, cout=128, k=8, stride=4, padding=2), 16, 500, 1),
    (dict(kind="convT", cin=128, cout=64, k=4, stride=2, padding=1), 16, 2000, 1),
    (dict(kind="convT", cin=64, cout=32, k=4, stride=2, padding=1), 16, 4000, 1),
    (dict(kind="conv", cin=160, cout=512, k=7, padding=3), 16, 100, 1),
]
for p, H in ((2, 4256), (11, 774)):
    h1 = (H + 4 - 5) // 3 + 1
    h2 = (h1 + 4 - 5) // 3 + 1
    h3 = (h2 + 4 - 5) // 3 + 1
    h4 = (h3 + 4 - 5) // 3 + 1
    SHAPES += [
        (dict(kind="conv", cin=32, cout=128, k=5, stride=3, padding=2), 32 * p, h1, p),
        (dict(kind="conv", cin=128, cout=512, k=5, stride=3, padding=2), 32 * p, h2, p),
        (dict(kind="conv", cin=512, cout=1024, k=5, stride=3, padding=2), 32 * p, h3, p),
        (dict(kind="conv", cin=1024, cout=1024, k=5, padding=2), 32 * p, h4, p),
    ]
SHAPES += [
    (dict(kind="conv", cin=128, cout=128, k=41, stride=4, padding=20, groups=4), 32, 8512, 1),
    (dict(kind="conv", cin=128, cout=256, k=41, stride=4, padding=20, groups=16), 32, 2128, 1),
    (dict(kind="conv", cin=256, cout=512, k=41, stride=4, padding=20, groups=16), 32, 532, 1),
    (dict(kind="conv", cin=512, cout=1024, k=41, stride=4, padding=20, groups=16), 32, 133, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=41, padding=20, groups=16), 32, 34, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=5, padding=2), 32, 34, 1),
]


def seq(N, L, C, ni):
    if ni == 1:
        return SeqT((torch.randn(N, L, C, device=DEV) * 0.5).to(torch.bfloat16), N, L, C)
    B = N // ni
    t = (torch.randn(B, L, ni, C, device=DEV) * 0.5).to(torch.bfloat16)
    return SeqT(t, N, L, C, n_inner=ni, s_outer=L * ni * C, s_inner=C, s_row=ni * C)


def timeit(fn, iters=10, reps=5):
    """GPU time per launch in us: `iters` launches captured in ONE CUDA graph (no host launch overhead
    between them, as in the train step), best of `reps` replays."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


NCU_SHAPES = [0, 2, 4, 6, 8, 10, 12, 14, 16, 24, 28, 29, 33, 34]   # indices into SHAPES for `ncu` mode


def ncu_mode():
    """One launch per direction of a few representative shapes (run under `ncu --set full`)."""
    _lib.load()
    for i in NCU_SHAPES:
        kw, N, lin, ni = SHAPES[i]
        spec = ConvSpec(**kw)
        lay = ConvLayer(spec, "l", BF16, BF16)
        w = torch.randn(spec.weight_shape(), device=DEV) * 0.05
        lay.bind({"l.weight": w, "l.bias": torch.zeros(spec.cout, device=DEV)})
        lay.prep()
        X = seq(N, lin, spec.cin, ni)
        Y = seq(N, spec.out_len(lin), spec.cout, ni)
        dX = X.like()
        grads = {"l.weight": torch.zeros_like(w), "l.bias": torch.zeros(spec.cout, device=DEV)}
        lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1)
        lay.dgrad(Y, dX=dX, mask=X, mask_slope=0.1)
        lay.wgrad(X, Y, grads)
        torch.cuda.synchronize()
        print(i, kw, N, lin, ni, flush=True)


def main():
    what = sys.argv[1:] or ["fwd", "dgrad", "wgrad"]
    if what == ["ncu"]:
        return ncu_mode()
    lib = _lib.load()
    sel = os.environ.get("SWEEP_SHAPES")          # e.g. "16-19,22-29": indices into SHAPES
    shapes = SHAPES
    if sel:
        idx = []
        for part in sel.split(","):
            a, _, b = part.partition("-")
            idx += list(range(int(a), int(b or a) + 1))
        shapes = [SHAPES[i] for i in idx]
    for kw, N, lin, ni in shapes:
        spec = ConvSpec(**kw)
        lay = ConvLayer(spec, "l", BF16, BF16)
        w = torch.randn(spec.weight_shape(), device=DEV) * 0.05
        lay.bind({"l.weight": w, "l.bias": torch.zeros(spec.cout, device=DEV)})
        lay.prep()
        lout = spec.out_len(lin)
        X = seq(N, lin, spec.cin, ni)
        Y = seq(N, lout, spec.cout, ni)
        dX = X.like()
        flops = 2.0 * N * (lout if spec.kind != "convT" else lin) * spec.cout * spec.cig * spec.k
        grads = {"l.weight": torch.zeros_like(w), "l.bias": torch.zeros(spec.cout, device=DEV)}
        runs = {"fwd": lambda: lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1),
                "dgrad": lambda: lay.dgrad(Y, dX=dX, mask=X, mask_slope=0.1),
                "wgrad": lambda: lay.wgrad(X, Y, grads)}
        print(f"== {kw} N={N} L={lin} ni={ni}  {flops / 1e9:.2f} GFLOP", flush=True)
        for name in what:
            fn = runs[name]
            try:
                lib.artic_debug_set(2, 0); lib.artic_debug_set(3, 0); lib.artic_debug_set(6, 0)
                base = timeit(fn)
                line = [f"auto {base:7.1f}us {flops / base / 1e6:6.1f}TF"]
                if name == "wgrad":
                    for bn in (32, 64, 128, 256):
                        if lay.kcog % bn:
                            continue
                        lib.artic_debug_set(6, bn)
                        line.append(f"bn{bn} {timeit(fn):6.1f}")
                else:
                    cog = lay.kcog if name == "fwd" else lay.kcig
                    for bn in (32, 64, 128, 256):
                        if cog % bn:
                            continue
                        for mt in (1, 2, 4):
                            lib.artic_debug_set(2, bn); lib.artic_debug_set(3, mt)
                            line.append(f"{bn}x{mt} {timeit(fn):6.1f}")
                print(f"   {name:5s} " + " | ".join(line), flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"   {name:5s} FAILED: {ex}", flush=True)
                if "launch" in str(ex) or "CUDA" in str(ex):
                    return
    lib.artic_debug_set(2, 0); lib.artic_debug_set(3, 0); lib.artic_debug_set(6, 0)


if __name__ == "__main__":
    main()
