#!/usr/bin/env python
"""Clock timeline of CTA 0 of the tcgen05 weight-gradient kernel for a few layer shapes (artic_debug_buffer): clock64
offsets from kernel start of: setup done (staging area zeroed, barriers, TMEM), first operand stage seen by the MMA warp,
last MMA committed, accumulator seen by the epilogue, epilogue (split-K reductions) done, CTA exit.
`python tools/wgrad_timeline.py`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import BF16  # noqa: E402
from articulatory_b200.convspec import ConvSpec  # noqa: E402
from articulatory_b200.engine import ConvLayer, SeqT  # noqa: E402

SHAPES = [
    (dict(kind="conv", cin=32, cout=32, k=11, dilation=5, padding=25), 16, 8000),
    (dict(kind="conv", cin=32, cout=32, k=11, dilation=1, padding=5), 16, 8000),
    (dict(kind="conv", cin=32, cout=32, k=7, dilation=1, padding=3), 16, 8000),
    (dict(kind="conv", cin=32, cout=32, k=3, dilation=1, padding=1), 16, 8000),
    (dict(kind="conv", cin=32, cout=32, k=3, dilation=5, padding=5), 16, 8000),
    (dict(kind="conv", cin=64, cout=64, k=3, dilation=1, padding=1), 16, 4000),
    (dict(kind="conv", cin=64, cout=64, k=3, dilation=3, padding=3), 16, 4000),
    (dict(kind="conv", cin=64, cout=64, k=7, dilation=3, padding=9), 16, 4000),
    (dict(kind="conv", cin=128, cout=128, k=3, dilation=1, padding=1), 16, 2000),
    (dict(kind="conv", cin=256, cout=256, k=7, dilation=1, padding=3), 16, 500),
    (dict(kind="conv", cin=1024, cout=1024, k=5, padding=2), 64, 53),
]
TAGS = [(2, "setup done"), (20, "first stage at the MMA warp"), (22, "last MMA committed"), (30, "accumulators at the epilogue"),
        (31, "reductions issued"), (32, "exit")]


def main():
    lib = _lib.load()
    dev = "cuda:0"
    buf = torch.zeros(8192, dtype=torch.int64, device=dev)
    for kw, N, L in SHAPES:
        spec = ConvSpec(**kw)
        lay = ConvLayer(spec, "l", BF16, BF16)
        lay.bind({"l.weight": torch.randn(spec.weight_shape(), device=dev) * 0.05, "l.bias": torch.zeros(spec.cout, device=dev)})
        lay.prep()
        X = SeqT(torch.randn(N, L, spec.cin, device=dev).bfloat16(), N, L, spec.cin)
        dY = SeqT(torch.randn(N, spec.out_len(L), spec.cout, device=dev).bfloat16(), N, spec.out_len(L), spec.cout)
        grads = {"l.weight": torch.zeros(spec.weight_shape(), device=dev), "l.bias": torch.zeros(spec.cout, device=dev)}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(3):
            lay.zero_wgrad()
            buf.zero_()
            lib.artic_debug_buffer(buf.data_ptr())
            e0.record()
            lay.wgrad(X, dY, grads)
            e1.record()
            torch.cuda.synchronize()
            lib.artic_debug_buffer(None)
        h = buf.cpu().tolist()
        t0 = h[201]
        line = ", ".join(f"{name} {h[200 + tag] - t0}" for tag, name in TAGS if h[200 + tag] > 0)
        print(f"{kw['cin']}->{kw['cout']} k{kw['k']} N={N} L={L}: launch {e0.elapsed_time(e1) * 1e3:.1f} us; CTA 0 clocks: {line}")


if __name__ == "__main__":
    main()
