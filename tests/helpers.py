"""Shared helpers for parity tests."""
import torch


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (a random-init G is
    near-silent, so element-wise relative error is meaningless)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def digest(t, n=64):
    f = t.detach().cpu().reshape(-1)
    stride = max(1, f.numel() // n)
    return dict(shape=tuple(t.shape), sum=float(f.double().sum()), abs_sum=float(f.double().abs().sum()),
                stride=stride, sample=f[::stride][:n].clone())


def check_digest(t, d, rtol, what="", sum_rtol=None):
    """Compare tensor ``t`` with a digest written by tests/golden/make_golden.py."""
    assert tuple(t.shape) == tuple(d["shape"]), f"{what}: shape {tuple(t.shape)} != {d['shape']}"
    g = digest(t)
    scale = max(d["abs_sum"], 1e-30)
    sum_rtol = rtol if sum_rtol is None else sum_rtol
    assert abs(g["abs_sum"] - d["abs_sum"]) <= sum_rtol * scale, f"{what}: abs_sum {g['abs_sum']} vs {d['abs_sum']}"
    if sum_rtol == rtol:
        assert abs(g["sum"] - d["sum"]) <= rtol * scale, f"{what}: sum {g['sum']} vs {d['sum']}"
    s_ref = d["sample"].double()
    err = (g["sample"].double() - s_ref).norm() / s_ref.norm().clamp_min(1e-30)
    assert err <= rtol, f"{what}: sample rel err {float(err)}"
