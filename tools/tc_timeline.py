#!/usr/bin/env python
"""Cycle timeline of CTA 0 of the tcgen05 conv kernel for a few shapes (artic_debug_buffer):
prints, per event tag, the clock64 offsets from kernel start.  Tags: 1 start, 2 setup done,
10 producer got A slot, 11 producer got W slot, 20 MMA saw A, 21 MMA saw W, 22 tile committed,
30 epilogue saw accumulator, 31 epilogue done."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import BF16  # noqa: E402
from articulatory_b200.convspec import ConvSpec  # noqa: E402
from articulatory_b200.engine import ConvLayer  # noqa: E402
from tools.tc_sweep import SHAPES, seq  # noqa: E402


def main():
    lib = _lib.load()
    buf = torch.zeros(8192, dtype=torch.int64, device="cuda:0")
    import os
    lib.artic_debug_set(9, int(os.environ.get('EPI_DBG', '0')))
    for i in (0, 2, 6, 12, 24):
        kw, N, lin, ni = SHAPES[i]
        spec = ConvSpec(**kw)
        lay = ConvLayer(spec, "l", BF16, BF16)
        w = torch.randn(spec.weight_shape(), device="cuda:0") * 0.05
        lay.bind({"l.weight": w, "l.bias": torch.zeros(spec.cout, device="cuda:0")})
        lay.prep()
        X = seq(N, lin, spec.cin, ni)
        Y = seq(N, spec.out_len(lin), spec.cout, ni)
        for rep in range(2):       # second run: warm L2 / descriptors
            buf.zero_()
            lib.artic_debug_buffer(buf.data_ptr())
            lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1)
            torch.cuda.synchronize()
            lib.artic_debug_buffer(None)
        h = buf.cpu().tolist()
        ev = {j: (h[1 + 2 * j], h[2 + 2 * j]) for j in range(40) if h[1 + 2 * j] > 0}
        t0 = ev[1][0]
        names = {1: "start", 2: "setup done", 20: "A seen (first/last kc)", 22: "tile committed (first/last)",
                 30: "epilogue saw acc (first/last)", 31: "epilogue done (first/last)",
                 32: "chunk prefetch issued", 33: "chunk tmem loaded", 34: "chunk staged", 35: "chunk stored"}
        print(f"== {kw} N={N} L={lin}")
        print("   " + "; ".join(f"{names.get(j, j)}: {a - t0}/{b - t0}" for j, (a, b) in sorted(ev.items())) + f"; exit: {h[90] - t0}")
        tr = [(h[3000 + 3 * j], h[3001 + 3 * j], h[3002 + 3 * j]) for j in range(48)]
        tr = [x for x in tr if x[0] > 0 and x[1] > 0]
        if tr:
            print("   W loads (issue, seen-by-MMA, mma-issued; latency): " +
                  " ".join(f"[{a - t0},{b - t0},{c - t0};{b - a}]" for a, b, c in tr))


if __name__ == "__main__":
    main()
