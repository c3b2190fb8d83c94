"""Stand-alone layer plugins on the tap-gather kernels.

``CausalConv1d`` / ``CausalConvTranspose1d`` mirror reference layers/causal_conv.py:12-66 (same
constructor keywords, same ``.conv`` / ``.deconv`` parameter names): a left-only padded
convolution is the ordinary tap-gather contraction with tap offsets ``j*dilation - (k-1)*dilation``
whose launch stops at the input length, and the trimmed transposed convolution is the polyphase
launch set cut ``stride`` samples short — no pad / slice tensors are materialised.
Tensors cross this boundary in the reference's (B, C, T) layout; CUDA only.
"""
import torch

from . import _lib
from .convspec import ConvSpec
from .engine import ConvLayer, SeqT


class _ConvFn(torch.autograd.Function):
    """y = layer(x) for one ConvLayer; x (B, C_in, T) -> (B, C_out, T_out), fp32 at the boundary."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        _lib.require_cuda(x, "x")
        lay = mod._layer()
        B, C, T = x.shape
        dt = _lib.TORCH_DTYPE[lay.in_code]
        X = SeqT(x.detach().permute(0, 2, 1).contiguous().to(dt), B, T, C)
        lo = lay.spec.out_len(T)
        Y = SeqT.empty(B, lo, lay.spec.cout, lay.out_code, x.device)
        lay.forward(X, Y=Y)
        ctx.mod, ctx.X, ctx.lay = mod, X, lay
        ctx.names = [n for n, _ in mod.named_parameters()]
        return Y.t.permute(0, 2, 1).float()

    @staticmethod
    def backward(ctx, dy):
        lay, X = ctx.lay, ctx.X
        B, Co, To = dy.shape
        dY = SeqT(dy.permute(0, 2, 1).contiguous().to(_lib.TORCH_DTYPE[lay.out_code]), B, To, Co)
        dX = X.like()
        lay.dgrad(dY, dX=dX)
        grads = {n: torch.zeros_like(p) for n, p in ctx.mod._engine_params().items()}
        lay.zero_wgrad()
        lay.wgrad(X, dY, grads)
        lay.finish_grads(grads)
        inv = {v: k for k, v in ctx.mod._name_map().items()}
        pg = tuple(grads.get(inv.get(n)) for n in ctx.names)
        return (None, dX.t.permute(0, 2, 1).float()) + pg


class _SingleConv(torch.nn.Module):
    """Shared plumbing: one torch parameter container + one ConvLayer bound to it."""
    _container = "conv"
    precision = "fp32"

    def _spec(self) -> ConvSpec:
        raise NotImplementedError

    def _name_map(self):
        """engine parameter name -> torch parameter name"""
        c = self._container
        out = {}
        for n, _ in getattr(self, c).named_parameters():
            out["l." + n] = f"{c}.{n}"
        return out

    def _engine_params(self):
        named = dict(self.named_parameters())
        return {k: named[v].data for k, v in self._name_map().items()}

    def _layer(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_lay", None) is None or self._key != key:
            code = _lib.BF16 if self.precision == "bf16" else _lib.F32
            lay = ConvLayer(self._spec(), "l", code, code, x3=self.precision == "bf16x3")
            lay.bind(self._engine_params())
            lay.prep()
            self._lay, self._key = lay, key
        return self._lay

    def forward(self, x):
        return _ConvFn.apply(self, x, *[p for _, p in self.named_parameters()])


class CausalConv1d(_SingleConv):
    """CausalConv1d (reference layers/causal_conv.py:12-42): (B, C_in, T) -> (B, C_out, T)."""
    _container = "conv"

    def __init__(self, in_channels, out_channels, kernel_size, dilation=1, bias=True, pad="ConstantPad1d",
                 pad_params={"value": 0.0}):
        super().__init__()
        if pad != "ConstantPad1d" or float(pad_params.get("value", 0.0)) != 0.0:
            raise NotImplementedError("only zero ConstantPad1d padding is implemented (the reference default)")
        self.conv = torch.nn.Conv1d(in_channels, out_channels, kernel_size, dilation=dilation, bias=bias)
        self._geom = (in_channels, out_channels, kernel_size, dilation)

    def _spec(self):
        ci, co, k, d = self._geom
        p = (k - 1) * d
        # symmetric padding p would give T + p outputs; the reference keeps the first T
        return ConvSpec("conv", ci, co, k=k, dilation=d, padding=p, trim_right=p)


class CausalConvTranspose1d(_SingleConv):
    """CausalConvTranspose1d (reference layers/causal_conv.py:45-66): deconv(x)[:, :, :-stride]."""
    _container = "deconv"

    def __init__(self, in_channels, out_channels, kernel_size, stride, bias=True):
        super().__init__()
        self.deconv = torch.nn.ConvTranspose1d(in_channels, out_channels, kernel_size, stride, bias=bias)
        self.stride = stride
        self._geom = (in_channels, out_channels, kernel_size, stride)

    def _spec(self):
        ci, co, k, s = self._geom
        return ConvSpec("convT", ci, co, k=k, stride=s, trim_right=s)
