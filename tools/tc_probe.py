#!/usr/bin/env python
"""Probe of the tcgen05 tap-gather kernel on a real B200: runs a matrix of conv shapes through
ConvLayer.forward / dgrad with the tensor-core path (both shared-memory-descriptor swizzle
phase modes) and with the CUDA-core kernel, and prints relative errors against torch fp64.
Not a test (prints, never asserts): `timeout 300 python tools/tc_probe.py`."""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from articulatory_b200 import _lib  # noqa: E402
from articulatory_b200._lib import BF16  # noqa: E402
from articulatory_b200.convspec import ConvSpec  # noqa: E402
from articulatory_b200.engine import ConvLayer, SeqT  # noqa: E402

DEV = "cuda:0"
CASES = [
    (dict(kind="conv", cin=64, cout=64, k=1, padding=0), 2, 256, 1),
    (dict(kind="conv", cin=64, cout=64, k=3, padding=1), 2, 256, 1),
    (dict(kind="conv", cin=64, cout=64, k=3, padding=1), 3, 500, 1),
    (dict(kind="conv", cin=128, cout=128, k=7, dilation=3, padding=9), 2, 2000, 1),
    (dict(kind="conv", cin=256, cout=256, k=11, dilation=5, padding=25), 2, 500, 1),
    (dict(kind="conv", cin=32, cout=32, k=11, dilation=5, padding=25), 2, 1000, 1),
    (dict(kind="conv", cin=32, cout=32, k=3, padding=1), 2, 8000, 1),
    (dict(kind="conv", cin=512, cout=1024, k=5, stride=1, padding=2), 6, 53, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=5, stride=1, padding=2), 14, 10, 7),
    (dict(kind="conv", cin=128, cout=512, k=5, stride=3, padding=2), 6, 158, 3),
    (dict(kind="convT", cin=512, cout=256, k=10, stride=5, padding=3, output_padding=1), 2, 100, 1),
    (dict(kind="convT", cin=64, cout=32, k=4, stride=2, padding=1), 2, 4000, 1),
    (dict(kind="conv", cin=96, cout=160, k=5, padding=2), 2, 300, 1),
    (dict(kind="conv", cin=64, cout=64, k=11, dilation=1, padding=5), 16, 4000, 1),
    (dict(kind="conv", cin=32, cout=32, k=7, dilation=3, padding=9), 16, 8000, 1),
    (dict(kind="conv", cin=256, cout=256, k=3, dilation=1, padding=1), 16, 500, 1),
    (dict(kind="conv", cin=128, cout=128, k=11, dilation=1, padding=5), 16, 2000, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=5, stride=1, padding=2), 32, 53, 2),
    (dict(kind="conv", cin=32, cout=32, k=41, stride=1, padding=20, groups=1), 2, 517, 1),
    (dict(kind="conv", cin=48, cout=96, k=3, padding=1), 2, 300, 1),
    (dict(kind="conv", cin=32, cout=128, k=5, stride=3, padding=2), 6, 473, 3),
    (dict(kind="conv", cin=512, cout=1024, k=5, stride=3, padding=2), 4, 158, 2),
    (dict(kind="conv", cin=512, cout=1024, k=5, stride=3, padding=2), 22, 29, 11),
    (dict(kind="conv", cin=128, cout=128, k=41, stride=4, padding=20, groups=4), 2, 8512, 1),
    (dict(kind="conv", cin=256, cout=512, k=41, stride=4, padding=20, groups=16), 2, 532, 1),
    (dict(kind="conv", cin=512, cout=1024, k=41, stride=4, padding=20, groups=16), 3, 133, 1),
    (dict(kind="conv", cin=1024, cout=1024, k=41, stride=1, padding=20, groups=16), 4, 34, 1),
    (dict(kind="conv", cin=64, cout=64, k=7, stride=2, dilation=3, padding=9), 3, 1001, 1),
]


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def torch_fwd(spec, x, w, b):
    if spec.kind == "conv":
        return F.conv1d(x, w, b, stride=spec.stride, padding=spec.padding, dilation=spec.dilation, groups=spec.groups)
    return F.conv_transpose1d(x, w, b, stride=spec.stride, padding=spec.padding, output_padding=spec.output_padding)


def seq_from(x, n_inner, code=BF16):
    """(N, C, L) fp64 -> SeqT in plain (N, L, C) or period (B, L, p, C) storage."""
    N, C, L = x.shape
    td = _lib.TORCH_DTYPE[code]
    if n_inner == 1:
        return SeqT(x.permute(0, 2, 1).contiguous().to(DEV, td), N, L, C)
    B = N // n_inner
    t = x.reshape(B, n_inner, C, L).permute(0, 3, 1, 2).contiguous().to(DEV, td)      # (B, L, p, C)
    return SeqT(t, N, L, C, n_inner=n_inner, s_outer=L * n_inner * C, s_inner=C, s_row=n_inner * C)


def seq_to(s: SeqT):
    if s.n_inner == 1:
        return s.t.float().cpu().permute(0, 2, 1)
    B = s.N // s.n_inner
    return s.t.float().cpu().permute(0, 2, 3, 1).reshape(s.N, s.C, s.L)


def run(case, mode):
    kw, N, lin, ni = case
    spec = ConvSpec(**kw)
    torch.manual_seed(1)
    w = torch.randn(spec.weight_shape(), dtype=torch.float64) / math.sqrt(spec.cig * spec.k)
    b = torch.randn(spec.cout, dtype=torch.float64) * 0.1
    x = torch.randn(N, spec.cin, lin, dtype=torch.float64).to(torch.bfloat16).double()
    w = w.to(torch.bfloat16).double()
    x.requires_grad_(True)
    x.requires_grad_(True)
    y = torch_fwd(spec, x, w, b)
    dy = torch.randn_like(y).to(torch.bfloat16).double()
    w.requires_grad_(True)
    y = torch_fwd(spec, x, w, b)
    gx, gw = torch.autograd.grad(y, [x, w], dy)
    w = w.detach()
    lib = _lib.load()
    lib.artic_debug_set(1, 1 if mode == "generic" else 0)
    lib.artic_debug_set(0, 1 if mode == "tc1" else 0)
    lay = ConvLayer(spec, "l", BF16, BF16)
    lay.bind({"l.weight": w.float().to(DEV).contiguous(), "l.bias": b.float().to(DEV)})
    lay.prep()
    X = seq_from(x.detach(), ni)
    lout = spec.out_len(lin)
    Y = X.like(C=spec.cout, L=lout)
    Y2 = X.like(C=spec.cout, L=lout)
    Y.t.fill_(7.0)
    lay.forward(X, Y=Y, Y2=Y2, act=_lib.ACT_LRELU, act_slope=0.1)
    torch.cuda.synchronize()
    e_f = rel(seq_to(Y), y.detach())
    e_a = rel(seq_to(Y2), F.leaky_relu(y.detach(), 0.1))
    dY = seq_from(dy, ni)
    dX = X.like()
    lay.dgrad(dY, dX=dX)
    torch.cuda.synchronize()
    e_d = rel(seq_to(dX), gx)
    params = {"l.weight": lay.v, "l.bias": lay.b}
    grads = {k: torch.zeros_like(t) for k, t in params.items()}
    lay.zero_wgrad()
    lay.wgrad(X, dY, grads)
    lay.finish_grads(grads)
    torch.cuda.synchronize()
    e_w = rel(grads["l.weight"].cpu(), gw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        lay.wgrad(X, dY, grads)
    e0.record()
    for _ in range(10):
        lay.wgrad(X, dY, grads)
    e1.record()
    torch.cuda.synchronize()
    ms_w = e0.elapsed_time(e1) / 10
    # timing of the forward
    for _ in range(3):
        lay.forward(X, Y=Y, Y2=Y2, act=_lib.ACT_LRELU, act_slope=0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lay.forward(X, Y=Y, Y2=Y2, act=_lib.ACT_LRELU, act_slope=0.1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = 2.0 * N * lout * spec.cout * spec.cig * spec.k / (spec.stride if spec.kind == "convT" else 1)
    return e_f, e_a, e_d, e_w, ms, flops / ms / 1e9, ms_w, flops / ms_w / 1e9


def main():
    modes = sys.argv[1:] or ["generic", "tc0", "tc1"]
    for ci, case in enumerate(CASES):
        for mode in modes:
            t0 = time.time()
            try:
                e_f, e_a, e_d, e_w, ms, tf, ms_w, tf_w = run(case, mode)
                print(f"case {ci:2d} {mode:8s} fwd {e_f:.2e} act {e_a:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e} | fwd {ms:.4f} ms {tf:7.2f} TF/s "
                      f"wgrad {ms_w:.4f} ms {tf_w:7.2f} TF/s  "
                      f"{case[0]} N={case[1]} L={case[2]} ni={case[3]}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"case {ci:2d} {mode:8s} FAILED after {time.time() - t0:.1f}s: {ex}", flush=True)
                if "CUDA" in str(ex) or "launch" in str(ex):
                    return


if __name__ == "__main__":
    main()
