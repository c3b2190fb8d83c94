"""Host-side index algebra of the tap-gather GEMM plans vs torch conv semantics (CPU).
Covers every conv shape family on the hot path: dilated 'same' convs (MRF), strided and
grouped convs (MSD), Conv2d(k,1) with stride 3 (MPD), k=2 pad=1 output conv, transposed
convs with the reference's padding rule (models/hifigan.py:82-103), Linear."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from articulatory_b200.convspec import ConvSpec
from tests import tap_emulator as E

CASES = [
    ConvSpec("conv", 6, 4, k=7, padding=3),
    ConvSpec("conv", 4, 4, k=3, dilation=5, padding=5),
    ConvSpec("conv", 4, 4, k=11, dilation=3, padding=15),
    ConvSpec("conv", 8, 8, k=41, stride=4, padding=20, groups=4),
    ConvSpec("conv", 8, 16, k=41, stride=4, padding=20, groups=4),
    ConvSpec("conv", 1, 4, k=15, padding=7),
    ConvSpec("conv", 4, 1, k=3, padding=1),
    ConvSpec("conv", 3, 5, k=5, stride=3, padding=2),
    ConvSpec("conv", 5, 1, k=2, stride=1, padding=1),
    ConvSpec("conv", 4, 4, k=3, stride=4, padding=1),        # phases without taps in dgrad
    ConvSpec("convT", 6, 4, k=10, stride=5, padding=3, output_padding=1),
    ConvSpec("convT", 4, 2, k=8, stride=4, padding=2),
    ConvSpec("convT", 4, 2, k=4, stride=2, padding=1),
    ConvSpec("convT", 3, 2, k=16, stride=8, padding=4),
    ConvSpec("convT", 4, 2, k=6, stride=3, padding=2, output_padding=1),   # egs/mri/voc1 (scales [8, 5, 3, 2])
    ConvSpec("linear", 7, 5),
]


def torch_forward(spec, x_ncl, w, b):
    if spec.kind == "conv":
        return F.conv1d(x_ncl, w, b, stride=spec.stride, padding=spec.padding, dilation=spec.dilation, groups=spec.groups)
    if spec.kind == "convT":
        return F.conv_transpose1d(x_ncl, w, b, stride=spec.stride, padding=spec.padding, output_padding=spec.output_padding)
    return F.linear(x_ncl.transpose(1, 2), w, b).transpose(1, 2)


@pytest.mark.parametrize("spec", CASES, ids=lambda s: f"{s.kind}-{s.cin}-{s.cout}-k{s.k}s{s.stride}d{s.dilation}g{s.groups}")
@pytest.mark.parametrize("lin", [1, 13, 40])
def test_plans_match_torch(spec, lin):
    if spec.kind == "linear" and lin != 1:
        pytest.skip("linear is used with a single row")
    if spec.out_len(lin) <= 0:
        pytest.skip("empty output")
    torch.manual_seed(0)
    w = torch.randn(spec.weight_shape(), dtype=torch.float64, requires_grad=True)
    b = torch.randn(spec.cout, dtype=torch.float64)
    x = torch.randn(2, spec.cin, lin, dtype=torch.float64, requires_grad=True)
    y = torch_forward(spec, x, w, b)
    lout = spec.out_len(lin)
    assert y.shape[2] == lout
    dy = torch.randn_like(y)
    dx_ref, dw_ref = torch.autograd.grad(y, [x, w], dy)

    xcl = x.detach().numpy().transpose(0, 2, 1)
    wf = E.prep_weight(w.detach().numpy(), spec, "fwd")
    wb = E.prep_weight(w.detach().numpy(), spec, "bwd")
    y_mine = E.tapconv(xcl, wf, spec.fwd_launches(lin), lout, b.numpy())
    assert np.allclose(y_mine.transpose(0, 2, 1), y.detach().numpy(), atol=1e-10)

    dycl = dy.numpy().transpose(0, 2, 1)
    dx_mine = E.tapconv(dycl, wb, spec.dgrad_launches(lin), lin)
    assert np.allclose(dx_mine.transpose(0, 2, 1), dx_ref.numpy(), atol=1e-10)

    A, B, *_ = spec.prep_strides("fwd")
    dwp = E.tapwgrad(xcl, dycl, spec.wgrad_launch(lin), spec.k, spec.groups, A, B)
    assert np.allclose(E.unprep_weight(dwp, spec), dw_ref.numpy(), atol=1e-10)


def test_taps_within_abi_limit():
    from articulatory_b200._lib import MAX_TAPS
    for spec in CASES:
        for L in spec.fwd_launches(64) + spec.dgrad_launches(64):
            assert 1 <= len(L.off) <= MAX_TAPS
