"""GPU parity of the individual C-ABI kernels against the CPU oracle (torch fp32/fp64 ops
that the reference itself calls).  Runs on the B200 box: `pytest -m gpu`."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from articulatory_b200 import _lib
    from articulatory_b200.convspec import ConvSpec
    from articulatory_b200.engine import ConvLayer, SeqT
    from articulatory_b200._lib import call, ptr, F32, BF16

DEV = "cuda:0"

CONV_CASES = [
    # (spec kwargs, N, Lin)
    (dict(kind="conv", cin=32, cout=32, k=3, dilation=1, padding=1), 3, 500),
    (dict(kind="conv", cin=64, cout=64, k=7, dilation=3, padding=9), 2, 333),
    (dict(kind="conv", cin=32, cout=32, k=11, dilation=5, padding=25), 2, 1000),
    (dict(kind="conv", cin=141, cout=96, k=7, padding=3), 2, 100),
    (dict(kind="conv", cin=1, cout=32, k=15, padding=7), 2, 2129),
    (dict(kind="conv", cin=32, cout=32, k=41, stride=4, padding=20, groups=4), 2, 517),
    (dict(kind="conv", cin=32, cout=64, k=41, stride=4, padding=20, groups=16), 2, 133),
    (dict(kind="conv", cin=64, cout=1, k=3, padding=1), 3, 34),
    (dict(kind="conv", cin=4, cout=16, k=5, stride=3, padding=2), 6, 406),
    (dict(kind="conv", cin=64, cout=64, k=5, stride=1, padding=2), 14, 10),
    (dict(kind="conv", cin=64, cout=1, k=2, stride=1, padding=1), 4, 16),
    (dict(kind="convT", cin=64, cout=32, k=10, stride=5, padding=3, output_padding=1), 2, 100),
    (dict(kind="convT", cin=32, cout=16, k=8, stride=4, padding=2), 2, 125),
    (dict(kind="convT", cin=16, cout=8, k=4, stride=2, padding=1), 3, 77),
    (dict(kind="linear", cin=512, cout=256), 5, 1),
]


def _torch_fwd(spec, x, w, b):
    if spec.kind == "conv":
        return F.conv1d(x, w, b, stride=spec.stride, padding=spec.padding, dilation=spec.dilation, groups=spec.groups)
    if spec.kind == "convT":
        return F.conv_transpose1d(x, w, b, stride=spec.stride, padding=spec.padding, output_padding=spec.output_padding)
    return F.linear(x.transpose(1, 2), w, b).transpose(1, 2)


@pytest.mark.parametrize("case", range(len(CONV_CASES)))
@pytest.mark.parametrize("wn", [False, True])
@pytest.mark.parametrize("prec", ["fp32", "bf16", "bf16x3"])
def test_conv_layer_fwd_dgrad_wgrad(case, wn, prec):
    kw, N, lin = CONV_CASES[case]
    spec = ConvSpec(**kw)
    if spec.kind == "linear" and wn:
        pytest.skip("no weight norm on Linear")
    code = BF16 if prec == "bf16" else F32
    # bf16x3: split-operand tensor-core contraction, ~2^-16 relative operand error (artic.h: artic_tapconv_t.X_sp)
    tol = {"fp32": 2e-5, "bf16": 2e-2, "bf16x3": 5e-5}[prec]
    torch.manual_seed(case)
    v = torch.randn(spec.weight_shape(), dtype=torch.float64) / math.sqrt(spec.cig * spec.k)
    g = (torch.rand(spec.weight_shape()[0], *([1] * (v.dim() - 1)), dtype=torch.float64) + 0.5) if wn else None
    b = torch.randn(spec.cout, dtype=torch.float64) * 0.1
    x = torch.randn(N, spec.cin, lin, dtype=torch.float64)
    v.requires_grad_(True)
    x.requires_grad_(True)
    if wn:
        g.requires_grad_(True)
        w = g * v / v.pow(2).sum(dim=tuple(range(1, v.dim())), keepdim=True).sqrt()
    else:
        w = v
    b.requires_grad_(True)
    y = _torch_fwd(spec, x, w, b)
    dy = torch.randn_like(y)
    ins = [x, v, b] + ([g] if wn else [])
    gr = torch.autograd.grad(y, ins, dy)

    lay = ConvLayer(spec, "l", code, code, x3=prec == "bf16x3")
    _lib.path_counts(reset=True)
    params = {"l.bias": b.detach().float().to(DEV)}
    if wn:
        params["l.weight_v"] = v.detach().float().to(DEV).contiguous()
        params["l.weight_g"] = g.detach().float().to(DEV).contiguous()
    else:
        params["l.weight"] = v.detach().float().to(DEV).contiguous()
    lay.bind(params)
    lay.prep()
    td = _lib.TORCH_DTYPE[code]
    X = SeqT(x.detach().permute(0, 2, 1).contiguous().to(DEV, td), N, lin, spec.cin)
    lout = spec.out_len(lin)
    Y = SeqT.empty(N, lout, spec.cout, code, DEV)
    Y2 = SeqT.empty(N, lout, spec.cout, code, DEV)
    lay.forward(X, Y=Y, Y2=Y2, act=_lib.ACT_LRELU, act_slope=0.1)
    torch.cuda.synchronize()
    y_mine = Y.t.float().cpu().permute(0, 2, 1)
    assert rel_err(y_mine, y) < tol
    assert rel_err(Y2.t.float().cpu().permute(0, 2, 1), F.leaky_relu(y, 0.1)) < tol

    dY = SeqT(dy.permute(0, 2, 1).contiguous().to(DEV, td), N, lout, spec.cout)
    dX = SeqT.empty(N, lin, spec.cin, code, DEV)
    lay.dgrad(dY, dX=dX)
    torch.cuda.synchronize()
    assert rel_err(dX.t.float().cpu().permute(0, 2, 1), gr[0]) < tol

    grads = {k: torch.zeros_like(t) for k, t in params.items()}
    lay.zero_wgrad()
    lay.wgrad(X, dY, grads)
    lay.finish_grads(grads)
    torch.cuda.synchronize()
    assert rel_err(grads["l.weight_v" if wn else "l.weight"].cpu(), gr[1]) < tol
    assert rel_err(grads["l.bias"].cpu(), gr[2]) < tol
    if wn:
        assert rel_err(grads["l.weight_g"].cpu(), gr[3]) < tol
    if prec == "bf16x3":
        pc = _lib.path_counts()
        if lay.x3:      # eligible shape: every contraction of the layer ran on the split-operand tcgen05 kernels
            assert pc["conv_tc_x3"] >= 2 and pc["conv_generic"] == 0 and pc["conv_c1"] == 0, pc
            assert pc["wgrad_tc_x3"] == 1 and pc["wgrad_generic"] == 0, pc
            assert pc["wgrad_bias_fused"] == (0 if lay.dw_swapped else 1), pc      # bias gradient from the same launch
            # the split copy written by the epilogue reproduces the fp32 output to ~2^-17
            sp = Y2.sp.float()
            assert rel_err((sp[0] + sp[1]).cpu(), Y2.t.cpu()) < 2e-5
        else:
            assert pc["conv_tc_x3"] == 0 and pc["wgrad_tc_x3"] == 0, pc


def test_epilogue_mask_res_alpha():
    """v = alpha*acc + bias; v += res_pre; v *= lrelu'(mask); v += res + res2."""
    spec = ConvSpec("conv", 16, 24, k=3, padding=1)
    torch.manual_seed(0)
    N, L = 2, 50
    w = torch.randn(spec.weight_shape()) * 0.2
    x = torch.randn(N, 16, L)
    res_pre, mask, res, res2 = (torch.randn(N, 24, L) for _ in range(4))
    acc = F.conv1d(x, w, None, padding=1)
    want = (0.25 * acc + res_pre) * torch.where(mask > 0, 1.0, 0.1) + res + res2
    lay = ConvLayer(spec, "l", F32, F32)
    lay.bind({"l.weight": w.to(DEV)})
    lay.prep()
    cl = lambda t: SeqT(t.permute(0, 2, 1).contiguous().to(DEV), N, L, t.shape[1])
    Y = SeqT.empty(N, L, 24, F32, DEV)
    lay.forward(cl(x), Y=Y, res_pre=cl(res_pre), mask=cl(mask), mask_slope=0.1, res=cl(res), res2=cl(res2), alpha=0.25)
    torch.cuda.synchronize()
    assert rel_err(Y.t.cpu().permute(0, 2, 1), want) < 1e-5


def test_period_layout_matches_conv2d():
    """n_inner = p addressing == the reference's view(b, c, t//p, p) + Conv2d((k,1),(s,1))."""
    p, B, H, Ci, Co = 3, 2, 40, 8, 12
    torch.manual_seed(1)
    x = torch.randn(B, Ci, H, p)
    w = torch.randn(Co, Ci, 5, 1) * 0.2
    b = torch.randn(Co) * 0.1
    want = F.conv2d(x, w, b, stride=(3, 1), padding=(2, 0))
    spec = ConvSpec("conv", Ci, Co, k=5, stride=3, padding=2)
    lay = ConvLayer(spec, "l", F32, F32)
    lay.bind({"l.weight": w.to(DEV), "l.bias": b.to(DEV)})
    lay.prep()
    X = SeqT.period(B, H, p, Ci, F32, DEV)
    X.t.copy_(x.permute(0, 2, 3, 1))
    Ho = spec.out_len(H)
    Y = SeqT.period(B, Ho, p, Co, F32, DEV)
    lay.forward(X, Y=Y)
    torch.cuda.synchronize()
    assert rel_err(Y.t.cpu().permute(0, 3, 1, 2), want) < 1e-5


def test_small_ops():
    torch.manual_seed(0)
    B, L = 3, 8512
    x = torch.randn(B, 1, L)
    xd = x.to(DEV)
    # avg pool fwd/bwd
    want = F.avg_pool1d(x, 4, 2, 2)
    lo = want.shape[2]
    y = torch.empty(B, lo, device=DEV)
    call("artic_avgpool1d", ptr(xd), ptr(y), B, L, lo, 4, 2, 2, F32)
    assert torch.allclose(y.cpu(), want[:, 0], atol=1e-6)
    xr = x.clone().requires_grad_(True)
    dy = torch.randn(B, 1, lo)
    F.avg_pool1d(xr, 4, 2, 2).backward(dy)
    dx = torch.zeros(B, L, device=DEV)
    dyd = dy.to(DEV)
    call("artic_avgpool1d_bwd", ptr(dyd), ptr(dx), B, L, lo, 4, 2, 2, 0, F32)
    assert torch.allclose(dx.cpu(), xr.grad[:, 0], atol=1e-6)
    # reflect pad fwd (bit exact) / bwd
    for pd in (2, 3, 7):
        want = F.pad(x, (0, pd), "reflect")
        yp = torch.empty(B, L + pd, device=DEV)
        call("artic_reflect_pad_right", ptr(xd), ptr(yp), B, L, L + pd, F32)
        assert torch.equal(yp.cpu(), want[:, 0])
        xr = x.clone().requires_grad_(True)
        dyp = torch.randn(B, 1, L + pd)
        F.pad(xr, (0, pd), "reflect").backward(dyp)
        dx = torch.zeros(B, L, device=DEV)
        dypd = dyp.to(DEV)
        call("artic_reflect_pad_right_bwd", ptr(dypd), ptr(dx), B, L, L + pd, 0, F32)
        assert torch.allclose(dx.cpu(), xr.grad[:, 0], atol=1e-6)
    # losses
    a, b = torch.randn(5, 7, 11), torch.randn(5, 7, 11)
    slot = torch.zeros(2, device=DEV)
    ad, bd = a.to(DEV), b.to(DEV)   # keep references: ptr() of a temporary would dangle
    call("artic_sqerr_sum", ptr(ad), a.numel(), 1.0, 1.0 / a.numel(), ptr(slot), F32)
    call("artic_l1_sum", ptr(ad), ptr(bd), a.numel(), 1.0 / a.numel(), ptr(slot[1:]), F32)
    assert abs(slot[0].item() - F.mse_loss(a, torch.ones_like(a)).item()) < 1e-5
    assert abs(slot[1].item() - F.l1_loss(a, b).item()) < 1e-5
    # adam vs torch.optim.Adam + MultiStepLR
    from articulatory_b200.optim import FusedAdam
    lin = torch.nn.Linear(13, 7).to(DEV)
    ref = torch.nn.Linear(13, 7)
    ref.load_state_dict({k: v.cpu() for k, v in lin.state_dict().items()})
    opt = FusedAdam(lin, lr=1e-3, betas=(0.5, 0.9), gamma=0.5, milestones=(3, 6))
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.5, 0.9))
    rs = torch.optim.lr_scheduler.MultiStepLR(ropt, gamma=0.5, milestones=[3, 6])
    for it in range(8):
        gw, gb = torch.randn(7, 13), torch.randn(7)
        opt.grad_views["weight"].copy_(gw)
        opt.grad_views["bias"].copy_(gb)
        ref.weight.grad, ref.bias.grad = gw.clone(), gb.clone()
        opt.step()
        ropt.step()
        rs.step()
        assert torch.allclose(lin.weight.detach().cpu(), ref.weight.detach(), atol=1e-6), it
    assert opt.step_count() == 8
    # every parameter starts on a 128-byte boundary of the flat buffers (the weight-side kernels' 16-byte row accesses)
    assert all(o % 32 == 0 for o in opt.offsets.values()) and opt.flat.numel() == 96 + 32
    assert all(v.data_ptr() % 128 == 0 for v in opt.views.values())


def test_adam_reads_the_bf16_wire_buffer():
    """artic_adam_step_wire with bf16 gradients (the data-parallel wire buffer, FusedAdam.wire) == torch.optim.Adam fed
    the same bf16-rounded gradients; sizes with a scalar tail and a long vector body."""
    from articulatory_b200.optim import FusedAdam
    torch.manual_seed(1)
    mod = torch.nn.Sequential(torch.nn.Linear(257, 129), torch.nn.Linear(129, 3)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(257, 129), torch.nn.Linear(129, 3))
    ref.load_state_dict({k: v.cpu() for k, v in mod.state_dict().items()})
    opt = FusedAdam(mod, lr=1e-3, betas=(0.5, 0.9), gamma=0.5, milestones=(2, 4))
    opt.wire = torch.zeros(opt.grad.numel(), dtype=torch.bfloat16, device=DEV)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.5, 0.9))
    rs = torch.optim.lr_scheduler.MultiStepLR(ropt, gamma=0.5, milestones=[2, 4])
    for it in range(5):
        opt.grad.fill_(float("nan"))                          # the fp32 gradient buffer must not be read
        for (n, p), rp in zip(mod.named_parameters(), ref.parameters()):
            g = torch.randn(p.shape).to(torch.bfloat16)
            o = opt.offsets[n]
            opt.wire[o:o + g.numel()].copy_(g.reshape(-1))
            rp.grad = g.float()
        opt.step()
        ropt.step()
        rs.step()
        for p, rp in zip(mod.parameters(), ref.parameters()):
            assert torch.allclose(p.detach().cpu(), rp.detach(), atol=1e-6), it


@pytest.mark.parametrize("res", [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240), (1024, 80, 1024)])
def test_stft_loss_single_resolution(res):
    from oracle import torch_oracle as O
    n_fft, hop, win = res
    torch.manual_seed(0)
    b = O.synthetic_batch(3, frames=25 if n_fft < 2048 else 100)
    T = b["y"].shape[2]
    y = b["y"][:, 0]
    x = (0.3 * torch.randn(3, T) + 0.5 * y).contiguous()
    window = torch.hann_window(win)
    sums = torch.zeros(3, device=DEV)
    xd, yd, wd = x.to(DEV), y.contiguous().to(DEV), window.to(DEV)
    call("artic_stft_loss_fwd", ptr(xd), ptr(yd), 3, T, n_fft, hop, win, ptr(wd), 1e-7, ptr(sums))
    x64 = x.double().requires_grad_(True)
    xm = O.stft_magnitude(x64, n_fft, hop, win, window.double())
    ym = O.stft_magnitude(y.double(), n_fft, hop, win, window.double())
    want = torch.stack([(ym - xm).pow(2).sum(), ym.pow(2).sum(), (ym.log() - xm.log()).abs().sum()])
    got = sums.cpu().double()
    assert torch.allclose(got, want.detach(), rtol=2e-4), (got, want)
    sc = torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")
    mag = F.l1_loss(torch.log(ym), torch.log(xm))
    (0.7 * sc + 1.3 * mag).backward()
    dx = torch.zeros(3, T, device=DEV)
    call("artic_stft_loss_bwd", ptr(xd), ptr(yd), 3, T, n_fft, hop, win, ptr(wd), 1e-7, ptr(sums), 0.7, 1.3, ptr(dx))
    # fp32 torch itself is ~5e-3 away from fp64 on this gradient (see tests/test_oracle_golden.py)
    x32 = x.clone().requires_grad_(True)
    xm32 = O.stft_magnitude(x32, n_fft, hop, win, window)
    ym32 = O.stft_magnitude(y, n_fft, hop, win, window)
    (0.7 * torch.norm(ym32 - xm32, p="fro") / torch.norm(ym32, p="fro") + 1.3 * F.l1_loss(ym32.log(), xm32.log())).backward()
    ref_noise = rel_err(x32.grad, x64.grad)
    assert rel_err(dx.cpu(), x64.grad) < max(1e-3, 3 * ref_noise), (rel_err(dx.cpu(), x64.grad), ref_noise)


def test_mr_stft_and_mel_modules(golden):
    from articulatory_b200.losses import MelSpectrogramLoss, MultiResolutionSTFTLoss
    from oracle import torch_oracle as O
    y_ = golden["g_out"].to(DEV).requires_grad_(True)
    y = golden["batch"]["y"].to(DEV)
    sc, mag = MultiResolutionSTFTLoss()(y_, y)
    (sc + mag).backward()
    assert abs(sc.item() - float(golden["stft"]["sc"])) < 1e-4 * float(golden["stft"]["sc"])
    assert abs(mag.item() - float(golden["stft"]["mag"])) < 1e-4 * float(golden["stft"]["mag"])
    assert rel_err(y_.grad.cpu(), golden["stft"]["grad"]) < 2e-2      # fp32-noise-limited, see above
    y_ = golden["g_out"].to(DEV).requires_grad_(True)
    ml = MelSpectrogramLoss(**O.E2W_MEL_LOSS_PARAMS).to(DEV)(y_, y)
    ml.backward()
    assert abs(ml.item() - float(golden["mel"]["loss"])) < 1e-4 * float(golden["mel"]["loss"])
    assert rel_err(y_.grad.cpu(), golden["mel"]["grad"]) < 1e-3


def test_mel_loss_and_grad_fused(golden):
    """artic_mel_loss_fwd_bwd (one launch, sparse filter ranges) == the golden loss and gradient of the
    reference MelSpectrogramLoss (tests/golden/make_golden.py), and the ranges cover every non-zero weight."""
    from articulatory_b200.losses import MelSpectrogramLoss
    from oracle import torch_oracle as O
    mod = MelSpectrogramLoss(**O.E2W_MEL_LOSS_PARAMS).to(DEV)
    mm = mod.melmat.cpu().numpy()                                  # (bins, mels)
    rng = mod.mel_ranges.cpu().numpy()
    n_bins, n_mels = mm.shape
    for m in range(n_mels):
        lo, hi = rng[2 * m], rng[2 * m + 1]
        assert not mm[:lo, m].any() and not mm[hi:, m].any()
    for k in range(n_bins):
        lo, hi = rng[2 * n_mels + 2 * k], rng[2 * n_mels + 2 * k + 1]
        assert not mm[k, :lo].any() and not mm[k, hi:].any()
    x = golden["g_out"].to(DEV).reshape(golden["g_out"].shape[0], -1).contiguous()
    y = golden["batch"]["y"].to(DEV).reshape(x.shape).contiguous()
    slot = torch.zeros(1, device=DEV)
    dx = torch.zeros_like(x)
    n = mod.numel(*x.shape)
    mod.loss_and_grad(x, y, 1.0 / n, slot, 1.0 / n, dx)
    torch.cuda.synchronize()
    assert abs(slot.item() - float(golden["mel"]["loss"])) < 1e-4 * float(golden["mel"]["loss"])
    assert rel_err(dx.cpu().view_as(golden["mel"]["grad"]), golden["mel"]["grad"]) < 1e-3


def test_tc_conv_with_unaligned_bias_and_views():
    """Parameters live in ONE flat buffer in the train step (FusedAdam), so bias pointers are only 4-byte
    aligned: the tensor-core epilogue must not assume 16-byte aligned bias (regression: misaligned address)."""
    torch.manual_seed(0)
    spec = ConvSpec(kind="conv", cin=64, cout=128, k=3, padding=1)
    w = torch.randn(spec.weight_shape()) * 0.05
    flat = torch.zeros(1 + spec.cout + 3, device=DEV)
    bias = flat[1:1 + spec.cout]                       # data_ptr % 16 == 4
    bias.copy_(torch.randn(spec.cout) * 0.1)
    assert bias.data_ptr() % 16 != 0
    x = torch.randn(3, spec.cin, 300)
    ref = F.leaky_relu(F.conv1d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), bias.cpu(), padding=1), 0.1)
    lay = ConvLayer(spec, "l", BF16, BF16)
    lay.bind({"l.weight": w.to(DEV).contiguous(), "l.bias": bias})
    lay.prep()
    X = SeqT(x.permute(0, 2, 1).contiguous().to(DEV, torch.bfloat16), 3, 300, spec.cin)
    Y = SeqT.empty(3, 300, spec.cout, BF16, DEV)
    lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1)
    torch.cuda.synchronize()
    assert rel_err(Y.t.float().cpu().permute(0, 2, 1), ref) < 2e-2


@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
@pytest.mark.parametrize("shape", [(256, 256, 7, 1, 4, 300), (512, 256, 3, 1, 6, 130), (256, 512, 5, 2, 2, 1000),
                                   (1024, 1024, 5, 1, 8, 53)])
def test_tc_conv_cluster_weight_multicast(prec, shape):
    """Deep layers stream their weights through a thread-block cluster (TMA multicast, tapconv_tc.cu Plan.cs): forward
    and data gradient against torch on the same operands, with the cluster path asserted taken (debug key 22 = 2 turns it
    on; it is off by default because it does not pay inside the train step) and bit-identical to the plain path."""
    cin, cout, k, dil, N, L = shape
    spec = ConvSpec(kind="conv", cin=cin, cout=cout, k=k, dilation=dil, padding=(k - 1) // 2 * dil)
    code = BF16 if prec == "bf16" else F32
    torch.manual_seed(cin + k)
    w = torch.randn(spec.weight_shape(), dtype=torch.float64) / math.sqrt(cin * k)
    b = torch.randn(cout, dtype=torch.float64) * 0.1
    x = torch.randn(N, cin, L, dtype=torch.float64)
    if prec == "bf16":
        x, w_eff = _bf16_round(x), _bf16_round(w)
    else:
        w_eff = w.float().double()
    x.requires_grad_(True)
    y = F.conv1d(x, w_eff, b, padding=spec.padding, dilation=dil)
    dy = torch.randn_like(y)
    if prec == "bf16":
        dy = _bf16_round(dy)
    gx, = torch.autograd.grad(y, [x], dy)
    lay = ConvLayer(spec, "l", code, code, x3=prec == "bf16x3")
    lay.bind({"l.weight": w.float().to(DEV).contiguous(), "l.bias": b.float().to(DEV)})
    lay.prep()
    td = _lib.TORCH_DTYPE[code]
    tol = 6e-3 if prec == "bf16" else 5e-5
    lib = _lib.load()
    outs = []
    for off in (0, 1):
        lib.artic_debug_set(22, 0 if off else 2)
        try:
            _lib.path_counts(reset=True)
            X = SeqT(x.detach().permute(0, 2, 1).contiguous().to(DEV, td), N, L, cin)
            Y = SeqT.empty(N, L, cout, code, DEV)
            lay.forward(X, Y=Y)
            dY = SeqT(dy.permute(0, 2, 1).contiguous().to(DEV, td), N, L, cout)
            dX = SeqT.empty(N, L, cin, code, DEV)
            lay.dgrad(dY, dX=dX)
            torch.cuda.synchronize()
            pc = _lib.path_counts()
        finally:
            lib.artic_debug_set(22, 0)
        assert pc["conv_tc_cluster"] == (0 if off else 2), pc
        assert rel_err(Y.t.float().cpu().permute(0, 2, 1), y.detach()) < tol
        assert rel_err(dX.t.float().cpu().permute(0, 2, 1), gx) < tol
        outs.append((Y.t.float().cpu(), dX.t.float().cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])     # same MMAs, same order


@pytest.mark.parametrize("k,pad,ni", [(3, 1, 1), (2, 1, 3)])
def test_logit_conv_data_gradient_channel1_kernel(k, pad, ni):
    """Data gradient of a C -> 1 logit conv in the bf16 mode (fp32 logit gradient in, bf16 feature gradient out) with the
    discriminator backward's epilogue: (dgrad + feature-matching gradient) x LeakyReLU'(saved activation).  Must run
    on the channel-1 kernel (it used to fall to the generic one), plain and period layouts."""
    C, B, H = 1024, 4, 23
    spec = ConvSpec(kind="conv", cin=C, cout=1, k=k, padding=pad)
    torch.manual_seed(k)
    w = torch.randn(spec.weight_shape(), dtype=torch.float64) * 0.05
    Ho = spec.out_len(H)
    dz = torch.randn(B * ni, 1, Ho, dtype=torch.float64)
    act = _bf16_round(torch.randn(B * ni, C, H))
    fm = _bf16_round(torch.randn(B * ni, C, H) * 0.1)
    gx = F.conv_transpose1d(dz, w.float().double(), padding=pad)          # dgrad of conv1d(stride 1)
    want = (gx + fm) * torch.where(act > 0, 1.0, 0.1)
    lay = ConvLayer(spec, "l", BF16, F32)
    lay.bind({"l.weight": w.float().to(DEV).contiguous()})
    lay.prep()

    def seqs(t, dtype, Cc, L):          # (B*ni, Cc, L) -> SeqT in plain or period storage
        if ni == 1:
            return SeqT(t.permute(0, 2, 1).contiguous().to(DEV, dtype), B, L, Cc)
        st = SeqT.period(B, L, ni, Cc, _lib.DTYPE_CODE[dtype], DEV)
        st.t.copy_(t.view(B, ni, Cc, L).permute(0, 3, 1, 2))
        return st

    dZ = seqs(dz, torch.float32, 1, Ho)
    A, FM = seqs(act, torch.bfloat16, C, H), seqs(fm, torch.bfloat16, C, H)
    dX = A.like()
    _lib.path_counts(reset=True)
    lay.dgrad(dZ, dX=dX, res_pre=FM, mask=A, mask_slope=0.1)
    torch.cuda.synchronize()
    pc = _lib.path_counts()
    assert pc["conv_c1"] == 1 and pc["conv_generic"] == 0, pc
    got = dX.t.float().cpu()
    got = got.permute(0, 2, 1) if ni == 1 else got.permute(0, 2, 3, 1).reshape(B * ni, C, H)
    assert rel_err(got, want) < 6e-3


@pytest.mark.parametrize("C,k,dil", [(32, 3, 1), (32, 7, 3), (32, 11, 5), (64, 3, 5), (64, 7, 1), (64, 7, 5), (32, 11, 1),
                                     (64, 11, 1), (64, 11, 3), (64, 11, 5)])
@pytest.mark.parametrize("N,L", [(3, 500), (2, 131), (5, 118)])
def test_fused_residual_unit(C, k, dil, N, L):
    """artic_resunit_fwd (conv1 dilated -> LeakyReLU -> conv2 -> + x, one tcgen05 launch, intermediate in shared memory)
    against torch on the same bf16 operands: the intermediate `at`, xn and axn; tile edges (L not a multiple of the
    118..126-row tiles; C = 64 with k = 11 runs with single-buffered tiles), zero padding of both convs at the sequence ends."""
    torch.manual_seed(C + k + dil)
    slope = 0.1
    specs = [ConvSpec(kind="conv", cin=C, cout=C, k=k, dilation=dil, padding=(k - 1) // 2 * dil),
             ConvSpec(kind="conv", cin=C, cout=C, k=k, padding=(k - 1) // 2)]
    ws = [torch.randn(sp.weight_shape(), dtype=torch.float64) / math.sqrt(C * k) for sp in specs]
    bs = [torch.randn(C, dtype=torch.float64) * 0.1 for _ in specs]
    x = _bf16_round(torch.randn(N, C, L))
    ax = _bf16_round(F.leaky_relu(x, slope))
    at = _bf16_round(F.leaky_relu(F.conv1d(ax, _bf16_round(ws[0]), bs[0].float().double(), padding=specs[0].padding, dilation=dil), slope))
    xn = F.conv1d(at, _bf16_round(ws[1]), bs[1].float().double(), padding=specs[1].padding) + x
    lays = []
    for i, sp in enumerate(specs):
        lay = ConvLayer(sp, f"l{i}", BF16, BF16)
        lay.bind({f"l{i}.weight": ws[i].float().to(DEV).contiguous(), f"l{i}.bias": bs[i].float().to(DEV)})
        lay.prep()
        lays.append(lay)
    cl = lambda t: t.permute(0, 2, 1).contiguous().to(DEV, torch.bfloat16)
    AX, X = cl(ax), cl(x)
    AT, Y, Y2 = (torch.full((N, L, C), float("nan"), dtype=torch.bfloat16, device=DEV) for _ in range(3))
    p = _lib.ResUnit()
    p.AX, p.XRES, p.W1t, p.W2t = ptr(AX), ptr(X), ptr(lays[0].Wb), ptr(lays[1].Wb)
    p.b1, p.b2 = ptr(lays[0].b), ptr(lays[1].b)
    p.AT, p.Y, p.Y2 = ptr(AT), ptr(Y), ptr(Y2)
    p.N, p.L, p.C, p.k, p.dil, p.slope = N, L, C, k, dil, slope
    call("artic_resunit_fwd", p)
    torch.cuda.synchronize()
    back = lambda t: t.float().cpu().permute(0, 2, 1)
    assert torch.isfinite(AT.float()).all() and torch.isfinite(Y.float()).all() and torch.isfinite(Y2.float()).all()
    assert rel_err(back(AT), at) < 6e-3
    assert rel_err(back(Y), xn) < 6e-3
    assert rel_err(back(Y2), F.leaky_relu(xn, slope)) < 6e-3


@pytest.mark.parametrize("C,k,dil", [(32, 3, 1), (32, 7, 3), (32, 11, 5), (64, 3, 5), (64, 7, 1), (64, 7, 5), (64, 11, 1), (64, 11, 5)])
@pytest.mark.parametrize("N,L", [(3, 500), (2, 131), (4, 78)])
def test_fused_residual_unit_data_gradient(C, k, dil, N, L):
    """artic_resunit_fwd mode 1: dt = conv2^T(gx) * lrelu'(at), gn = conv1^T(dt) * lrelu'(ax) + gx against torch autograd
    of the same unit on the same bf16 operands (dt is bf16-rounded before the second transposed conv, as in the kernel)."""
    torch.manual_seed(7 * C + k + dil)
    slope = 0.1
    specs = [ConvSpec(kind="conv", cin=C, cout=C, k=k, dilation=dil, padding=(k - 1) // 2 * dil),
             ConvSpec(kind="conv", cin=C, cout=C, k=k, padding=(k - 1) // 2)]
    ws = [_bf16_round(torch.randn(sp.weight_shape(), dtype=torch.float64) / math.sqrt(C * k)) for sp in specs]
    ax = _bf16_round(torch.randn(N, C, L))                   # saved activations: only their signs matter here
    at = _bf16_round(torch.randn(N, C, L))
    gx = _bf16_round(torch.randn(N, C, L))
    dt = _bf16_round(F.conv_transpose1d(gx, ws[1], padding=specs[1].padding) * torch.where(at > 0, 1.0, slope))
    gn = F.conv_transpose1d(dt, ws[0], padding=specs[0].padding, dilation=dil) * torch.where(ax > 0, 1.0, slope) + gx
    lays = []
    for i, sp in enumerate(specs):
        lay = ConvLayer(sp, f"l{i}", BF16, BF16)
        lay.bind({f"l{i}.weight": ws[i].float().to(DEV).contiguous()})
        lay.prep()
        lays.append(lay)
    cl = lambda t: t.permute(0, 2, 1).contiguous().to(DEV, torch.bfloat16)
    GX, AT, AX = cl(gx), cl(at), cl(ax)
    DT, GN = (torch.full((N, L, C), float("nan"), dtype=torch.bfloat16, device=DEV) for _ in range(2))
    p = _lib.ResUnit()
    p.AX = p.XRES = ptr(GX)
    p.W1t, p.W2t = ptr(lays[1].Wf), ptr(lays[0].Wf)
    p.M1, p.M2, p.AT, p.Y = ptr(AT), ptr(AX), ptr(DT), ptr(GN)
    p.N, p.L, p.C, p.k, p.dil, p.slope, p.mode = N, L, C, k, dil, slope, 1
    call("artic_resunit_fwd", p)
    torch.cuda.synchronize()
    back = lambda t: t.float().cpu().permute(0, 2, 1)
    assert torch.isfinite(DT.float()).all() and torch.isfinite(GN.float()).all()
    assert rel_err(back(DT), dt) < 6e-3
    assert rel_err(back(GN), gn) < 6e-3


def test_fused_residual_unit_refuses_what_does_not_fit():
    """C = 128: the weights cannot stay resident -> ARTIC_ENOSUP (the engine issues the two convs separately)."""
    p = _lib.ResUnit()
    t = torch.zeros(128 * 128 * 3, dtype=torch.bfloat16, device=DEV)
    p.AX = p.XRES = p.W1t = p.W2t = p.Y = ptr(t)
    p.N, p.L, p.C, p.k, p.dil, p.slope = 1, 64, 128, 3, 1, 0.1
    with pytest.raises(_lib.ArticError):
        call("artic_resunit_fwd", p)


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float64)


WEIGHT_CASES = [
    dict(kind="conv", cin=1024, cout=1024, k=5, stride=1, padding=2),                    # the discriminators' big layers
    dict(kind="conv", cin=1024, cout=1024, k=41, stride=1, padding=20, groups=16),       # K = 41: 8 inner columns per tile
    dict(kind="conv", cin=128, cout=256, k=41, stride=4, padding=20, groups=16),         # merged narrow groups (8 -> 16 per group)
    dict(kind="conv", cin=128, cout=128, k=11, dilation=5, padding=25),                  # K = 11: 352-float rows
    dict(kind="conv", cin=141, cout=256, k=7, padding=3),                                # padded input width
    dict(kind="conv", cin=1, cout=128, k=15, padding=7),                                 # one input channel
    dict(kind="conv", cin=32, cout=1, k=7, padding=3),                                   # one output channel
    dict(kind="conv", cin=40, cout=72, k=3, padding=1),                                  # ragged tiles
    dict(kind="convT", cin=256, cout=128, k=10, stride=5, padding=3, output_padding=1),  # outer = in-ch, even K
    dict(kind="convT", cin=64, cout=32, k=4, stride=2, padding=1),
    dict(kind="linear", cin=512, cout=256),
]


@pytest.mark.parametrize("case", range(len(WEIGHT_CASES)))
@pytest.mark.parametrize("wn", [False, True])
@pytest.mark.parametrize("prec", ["bf16", "bf16x3", "fp32"])
def test_row_run_weight_kernels_match_generic_tile_kernels(case, wn, prec, monkeypatch):
    """artic_weights_prep / _unprep: the row-run kernels (default) against the generic 32 x 32 x 8-tap tile kernels
    (ARTIC_WEIGHTS_GENERIC) on the same parameters — prepared weights BIT-identical in both layouts, gradients of v / g
    equal up to the summation order of the weight-norm dot products — and against torch for the weight itself."""
    from articulatory_b200 import engine
    spec = ConvSpec(**WEIGHT_CASES[case])
    if spec.kind == "linear" and wn:
        pytest.skip("no weight norm on Linear")
    code = BF16 if prec == "bf16" else F32
    torch.manual_seed(100 + case)
    v = torch.randn(spec.weight_shape()) / math.sqrt(spec.cig * spec.k)
    params = {"l.bias": torch.randn(spec.cout).to(DEV)}
    if wn:
        params["l.weight_v"] = v.to(DEV).contiguous()
        params["l.weight_g"] = (torch.rand(spec.weight_shape()[0], *([1] * (v.dim() - 1))) + 0.5).to(DEV).contiguous()
    else:
        params["l.weight"] = v.to(DEV).contiguous()
    out = {}
    for generic in (True, False):
        monkeypatch.setattr(engine, "_WEIGHTS_GENERIC", generic)
        lay = ConvLayer(spec, "l", code, code, pad_in=True, x3=prec == "bf16x3")
        lay.bind(params)
        lay.prep()
        ws = lay._single()
        assert (ws.total_tiles2 == 0) == generic and (ws.total_tiles == 0) != generic
        g = torch.Generator(device="cpu").manual_seed(7)
        lay.dWf.copy_(torch.randn(lay.dWf.shape, generator=g).to(DEV))
        grads = {k: torch.full_like(t, float("nan")) for k, t in params.items()}
        lay.finish_grads(grads)
        torch.cuda.synchronize()
        out[generic] = (lay.Wf.clone(), lay.Wb.clone(), {k: t.clone() for k, t in grads.items() if "weight" in k})
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    for k in out[True][2]:
        a, b = out[True][2][k], out[False][2][k]
        assert torch.isfinite(b).all(), k
        assert rel_err(b.cpu(), a.cpu()) < 1e-5, k
    # the prepared 'fwd' weight against torch: w = g * v / ||v|| in the layout [K][G/m][a_pad][b_pad]
    if wn:
        gg = params["l.weight_g"].cpu().double()
        vv = v.double()
        w = gg * vv / vv.pow(2).sum(dim=tuple(range(1, vv.dim())), keepdim=True).sqrt()
    else:
        w = v.double()
    lay_mg = lay.mg
    if lay_mg == 1 and lay.kcig == spec.cig:
        if spec.kind == "conv":       # [cout][cig][K] -> [K][G][cig][cog]
            ref = w.view(spec.groups, spec.cog, spec.cig, spec.k).permute(3, 0, 2, 1)
        elif spec.kind == "convT":    # [cin][cout][K] -> [K][1][cin][cout]
            ref = w.permute(2, 0, 1).unsqueeze(1)
        else:                         # [cout][cin] -> [1][1][cin][cout]
            ref = w.t().reshape(1, 1, spec.cin, spec.cout)
        assert rel_err(out[False][0].double().cpu(), ref) < (4e-3 if prec == "bf16" else 1e-6)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("B,hidden", [(1, 256), (16, 256), (33, 256), (5, 100)])
def test_fused_past_fc_encoder(dtype, B, hidden):
    """artic_mlp_fwd (the PastFCEncoder, 5 Linear layers with LeakyReLU(0.1) between, one launch) against torch float64 on
    the same (storage-rounded) weights; every saved activation is compared, not only the result.  hidden = 100 is a
    width the one-launch kernel does not take: the same call then runs the layers one by one."""
    from articulatory_b200.engine import mlp_forward
    torch.manual_seed(B)
    code = F32 if dtype == "f32" else BF16
    rnd = (lambda t: t.double()) if dtype == "f32" else _bf16_round
    dims = [512, hidden, hidden, hidden, hidden, 128]
    ws = [torch.randn(dims[i + 1], dims[i]) / math.sqrt(dims[i]) for i in range(5)]
    bs = [torch.randn(dims[i + 1]) * 0.1 for i in range(5)]
    x = torch.randn(B, 512)
    lays = []
    for i in range(5):
        lay = ConvLayer(ConvSpec("linear", dims[i], dims[i + 1]), f"m{i}", code, code)
        lay.bind({f"m{i}.weight": ws[i].to(DEV).contiguous(), f"m{i}.bias": bs[i].to(DEV)})
        lay.prep()
        lays.append(lay)
    acts = [SeqT.empty(B, 1, d, code, DEV) for d in dims]
    for a in acts:
        a.t.fill_(float("nan"))
    launches = _lib.launch_count
    mlp_forward(x.to(DEV), lays, acts, code, 0.1)
    torch.cuda.synchronize()
    assert _lib.launch_count - launches == (1 if hidden == 256 else 6)
    h = rnd(x)
    tol = 1e-5 if dtype == "f32" else 6e-3
    assert rel_err(acts[0].t.float().cpu().reshape(B, -1), h) < tol
    for i in range(5):
        h = F.linear(h, rnd(ws[i]), bs[i].double())
        if i < 4:
            h = F.leaky_relu(h, 0.1)
        h = rnd(h)
        got = acts[i + 1].t.float().cpu().reshape(B, -1)
        assert torch.isfinite(got).all()
        assert rel_err(got, h) < tol, i


MIXED_CASES = [
    # (spec kwargs, N, Lin, in dtype): the shapes whose bf16-mode kernels the uniform-precision cases above do not reach
    (dict(kind="conv", cin=1, cout=128, k=15, padding=7), 3, 1500, "f32"),               # Cin = 1 kernels (MSD first layer)
    (dict(kind="conv", cin=1, cout=32, k=5, stride=3, padding=2), 4, 901, "f32"),        # Cin = 1, strided (MPD first layer)
    (dict(kind="conv", cin=128, cout=256, k=5, stride=3, padding=2), 4, 301, "bf16"),    # strided tcgen05 wgrad (per-phase groups)
    (dict(kind="conv", cin=256, cout=256, k=5, stride=1, padding=2), 3, 200, "bf16"),
    (dict(kind="convT", cin=128, cout=64, k=8, stride=4, padding=2), 2, 200, "bf16"),    # transposed: wgrad with X / dY exchanged
    (dict(kind="conv", cin=141, cout=256, k=7, padding=3), 2, 100, "bf16"),              # padded input width (141 -> 256)
]


@pytest.mark.parametrize("case", range(len(MIXED_CASES)))
def test_bf16_mode_layers_fwd_dgrad_wgrad(case):
    """The bf16-mode kernels of the discriminator's first layers (fp32 signal in, bf16 features out: register-
    weight Cin = 1 forward, shared-memory-staged Cout = 1 data gradient into fp32, vector Cin = 1 weight gradient)
    and of the strided / transposed / padded tensor-core layers, against torch on the SAME rounded operands
    (so only the accumulation order differs)."""
    kw, N, lin, in_dt = MIXED_CASES[case]
    spec = ConvSpec(**kw)
    in_code = F32 if in_dt == "f32" else BF16
    torch.manual_seed(100 + case)
    w = torch.randn(spec.weight_shape(), dtype=torch.float64) / math.sqrt(spec.cig * spec.k)
    b = torch.randn(spec.cout, dtype=torch.float64) * 0.1
    x = torch.randn(N, spec.cin, lin, dtype=torch.float64)
    if in_code == BF16:
        x, w_eff = _bf16_round(x), _bf16_round(w)          # the layer computes with bf16 weights and activations
    else:
        w_eff = w.float().double()                          # Cin = 1 layers keep fp32 weights
    x.requires_grad_(True)
    w_eff.requires_grad_(True)
    b.requires_grad_(True)
    y = _torch_fwd(spec, x, w_eff, b)
    dy = _bf16_round(torch.randn_like(y))
    gx, gw, gb = torch.autograd.grad(y, [x, w_eff, b], dy)

    lay = ConvLayer(spec, "l", in_code, BF16, pad_in=True)
    params = {"l.weight": w.float().to(DEV).contiguous(), "l.bias": b.detach().float().to(DEV)}
    lay.bind(params)
    lay.prep()
    xin = x.detach().permute(0, 2, 1).contiguous()
    if spec.groups == 1 and lay.kcig > spec.cin:
        xin = F.pad(xin, (0, lay.kcig - spec.cin))          # zero-padded channels (the generator's input assembly does this)
    X = SeqT(xin.to(DEV, _lib.TORCH_DTYPE[in_code]), N, lin, xin.shape[2])
    lout = spec.out_len(lin)
    Y = SeqT.empty(N, lout, spec.cout, BF16, DEV)
    lay.forward(X, Y2=Y, act=_lib.ACT_LRELU, act_slope=0.1)
    torch.cuda.synchronize()
    assert rel_err(Y.t.float().cpu().permute(0, 2, 1), F.leaky_relu(y.detach(), 0.1)) < 6e-3      # bf16 output rounding

    dY = SeqT(dy.permute(0, 2, 1).contiguous().to(DEV, torch.bfloat16), N, lout, spec.cout)
    dX = SeqT.empty(N, lin, xin.shape[2], F32 if in_code == F32 else BF16, DEV)
    lay.dgrad(dY, dX=dX)
    torch.cuda.synchronize()
    got = dX.t.float().cpu()[:, :, :spec.cin].permute(0, 2, 1)
    assert rel_err(got, gx) < 6e-3          # bf16 'bwd' weights (also for the fp32-input layers) / bf16 dX

    grads = {k: torch.zeros_like(t) for k, t in params.items()}
    lay.zero_wgrad()
    lay.wgrad(X, dY, grads)
    lay.finish_grads(grads)
    torch.cuda.synchronize()
    assert rel_err(grads["l.weight"].cpu(), gw) < 2e-3
    assert rel_err(grads["l.bias"].cpu(), gb) < 2e-3
