"""GeneratorAdversarialLoss / DiscriminatorAdversarialLoss / FeatureMatchLoss on the fused
reduction kernels (artic_sqerr_sum, artic_l1_sum and their backward).

Same constructor arguments and call signatures as reference
articulatory/losses/adversarial_loss.py:12-123 and feat_match_loss.py:12-54.
Only loss_type="mse" (the reference default, used by every shipped config) is implemented.
"""
import torch

from .. import _lib
from .._lib import DTYPE_CODE, call, ptr


def _dense(t, order=None):
    """View ``t`` in its storage order (no copy for the permuted channels-last views the
    discriminator returns).  Returns (dense contiguous tensor, dim order used)."""
    _lib.require_cuda(t, "loss input")
    if order is None:
        order = sorted(range(t.dim()), key=lambda d: (-t.stride(d), d))
    tp = t.permute(order)
    if not tp.is_contiguous():
        tp = tp.contiguous()
    if tp.dtype not in DTYPE_CODE:
        tp = tp.float()
    return tp, order


def _inverse(order):
    inv = [0] * len(order)
    for i, d in enumerate(order):
        inv[d] = i
    return inv


class _SqErrMean(torch.autograd.Function):
    """mean((x - target)^2) — F.mse_loss against a constant tensor."""

    @staticmethod
    def forward(ctx, x, target):
        xd, order = _dense(x)
        slot = torch.zeros(1, dtype=torch.float32, device=x.device)
        call("artic_sqerr_sum", ptr(xd), xd.numel(), float(target), 1.0 / xd.numel(), ptr(slot), DTYPE_CODE[xd.dtype])
        ctx.save_for_backward(xd)
        ctx.order, ctx.target, ctx.in_dtype = order, float(target), x.dtype
        return slot[0]

    @staticmethod
    def backward(ctx, g):
        (xd,) = ctx.saved_tensors
        dx = torch.empty_like(xd)
        call("artic_sqerr_bwd", ptr(xd), xd.numel(), ctx.target, float(g) / xd.numel(), ptr(dx), 0, DTYPE_CODE[xd.dtype])
        return dx.permute(_inverse(ctx.order)).to(ctx.in_dtype), None


class _L1Mean(torch.autograd.Function):
    """mean(|a - b|) with gradient to ``a`` only (b is detached by the reference)."""

    @staticmethod
    def forward(ctx, a, b):
        ad, order = _dense(a)
        bd, _ = _dense(b.detach(), order)
        if bd.dtype != ad.dtype:
            bd = bd.to(ad.dtype)
        slot = torch.zeros(1, dtype=torch.float32, device=a.device)
        call("artic_l1_sum", ptr(ad), ptr(bd), ad.numel(), 1.0 / ad.numel(), ptr(slot), DTYPE_CODE[ad.dtype])
        ctx.save_for_backward(ad, bd)
        ctx.order, ctx.in_dtype = order, a.dtype
        return slot[0]

    @staticmethod
    def backward(ctx, g):
        ad, bd = ctx.saved_tensors
        da = torch.empty_like(ad)
        call("artic_l1_bwd", ptr(ad), ptr(bd), ad.numel(), float(g) / ad.numel(), ptr(da), 0, DTYPE_CODE[ad.dtype])
        return da.permute(_inverse(ctx.order)).to(ctx.in_dtype), None


def _check_type(loss_type):
    assert loss_type in ["mse", "hinge"], f"{loss_type} is not supported."
    if loss_type != "mse":
        raise NotImplementedError("only loss_type='mse' is implemented on the B200 path")


class GeneratorAdversarialLoss(torch.nn.Module):
    def __init__(self, average_by_discriminators=True, loss_type="mse"):
        super().__init__()
        _check_type(loss_type)
        self.average_by_discriminators = average_by_discriminators

    def forward(self, outputs):
        if isinstance(outputs, (tuple, list)):
            adv_loss = 0.0
            for i, outputs_ in enumerate(outputs):
                if isinstance(outputs_, (tuple, list)):
                    outputs_ = outputs_[-1]
                adv_loss = adv_loss + _SqErrMean.apply(outputs_, 1.0)
            if self.average_by_discriminators:
                adv_loss = adv_loss / (i + 1)
            return adv_loss
        return _SqErrMean.apply(outputs, 1.0)


class DiscriminatorAdversarialLoss(torch.nn.Module):
    def __init__(self, average_by_discriminators=True, loss_type="mse"):
        super().__init__()
        _check_type(loss_type)
        self.average_by_discriminators = average_by_discriminators

    def forward(self, outputs_hat, outputs):
        if isinstance(outputs, (tuple, list)):
            real_loss, fake_loss = 0.0, 0.0
            for i, (oh, o) in enumerate(zip(outputs_hat, outputs)):
                if isinstance(oh, (tuple, list)):
                    oh, o = oh[-1], o[-1]
                real_loss = real_loss + _SqErrMean.apply(o, 1.0)
                fake_loss = fake_loss + _SqErrMean.apply(oh, 0.0)
            if self.average_by_discriminators:
                fake_loss = fake_loss / (i + 1)
                real_loss = real_loss / (i + 1)
            return real_loss, fake_loss
        return _SqErrMean.apply(outputs, 1.0), _SqErrMean.apply(outputs_hat, 0.0)


class FeatureMatchLoss(torch.nn.Module):
    def __init__(self, average_by_layers=True, average_by_discriminators=True, include_final_outputs=False):
        super().__init__()
        self.average_by_layers = average_by_layers
        self.average_by_discriminators = average_by_discriminators
        self.include_final_outputs = include_final_outputs

    def forward(self, feats_hat, feats):
        total = 0.0
        for i, (fh, f) in enumerate(zip(feats_hat, feats)):
            li = 0.0
            if not self.include_final_outputs:
                fh, f = fh[:-1], f[:-1]
            for j, (a, b) in enumerate(zip(fh, f)):
                li = li + _L1Mean.apply(a, b)
            if self.average_by_layers:
                li = li / (j + 1)
            total = total + li
        if self.average_by_discriminators:
            total = total / (i + 1)
        return total
