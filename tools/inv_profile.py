#!/usr/bin/env python
"""Stage times of the inversion encoder forward (BASELINE configs[4] shape) with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import F32, call, ptr  # noqa: E402
from articulatory_b200.engine import SeqT  # noqa: E402
from articulatory_b200.models import BiGRU  # noqa: E402

dev = torch.device("cuda", 0)
N, C, T, H = 128, 1024, 400, 256
torch.manual_seed(0)
m = BiGRU(in_channels=C, hidden_size=H, out_channels=12).eval().to(dev)
P = m._prepare()
x = SeqT(torch.randn(N, T, C, device=dev), N, T, C)
gi = SeqT.empty(N, T, 6 * H, F32, dev)
h1 = SeqT.empty(N, T, 2 * H, F32, dev)
y = SeqT.empty(N, T, 12, F32, dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("split x      %.3f ms" % timeit(lambda: (x._h.__setitem__(0, None), x.ensure_split())))
print("gi1 (1024->1536, x3 tc)  %.3f ms" % timeit(lambda: P["gru1"]["lin"].forward(x, Y=gi, sp_y2=False)))
for Tn in (400, 100):
    print("gru layer T=%d  %.3f ms" % (Tn, timeit(lambda: call("artic_bigru_layer", ptr(gi.t), ptr(P["gru1"]["w_hh"]),
                                                                 ptr(P["gru1"]["b_hh"]), ptr(h1.t), N, Tn, H))))
print("gru layer N=16 T=400  %.3f ms" % timeit(lambda: call("artic_bigru_layer", ptr(gi.t), ptr(P["gru1"]["w_hh"]),
                                                            ptr(P["gru1"]["b_hh"]), ptr(h1.t), 16, T, H)))
print("gi2 (512->1536)  %.3f ms" % timeit(lambda: P["gru2"]["lin"].forward(h1, Y=gi, sp_y2=False)))
print("head (512->12)  %.3f ms" % timeit(lambda: P["head"]["lin"].forward(h1, Y=y)))
print(_lib.path_counts())
