"""MultiResolutionSTFTLoss / MelSpectrogramLoss on the fused sm_100a spectral kernels.

Mirrors reference articulatory/losses/stft_loss.py:121-170 and mel_loss.py:114-166 (same
constructor arguments, same forward signatures and return values).  The forward launches
one kernel per resolution that frames, FFTs (shared memory) and reduces; the backward
recomputes the spectrogram and overlap-adds dL/dx.  No spectrogram is materialised.
"""
import math

import numpy as np
import torch

from .. import _lib
from .._lib import call, ptr
from ..mel import slaney_mel_basis


def _as_2d(x):
    _lib.require_cuda(x, "signal")
    if x.dim() == 3:
        x = x.reshape(-1, x.size(2))
    return x.contiguous().float()


class STFTResolution:
    """One (fft_size, hop, win_length, window) resolution; holds the window on the device."""

    def __init__(self, fft_size, hop_size, win_length, window="hann_window"):
        self.fft_size, self.hop, self.win_length = fft_size, hop_size, win_length
        self.window_cpu = getattr(torch, window)(win_length).float()
        self._win = {}

    def window(self, device):
        w = self._win.get(device)
        if w is None:
            w = self._win[device] = self.window_cpu.to(device)
        return w

    def numel(self, B, T):
        return B * (1 + T // self.hop) * (self.fft_size // 2 + 1)

    def forward(self, x, y, sums):
        """sums (3,) fp32 device, pre-zeroed: += [sum (Y-X)^2, sum Y^2, sum |ln Y - ln X|]."""
        B, T = x.shape
        call("artic_stft_loss_fwd", ptr(x), ptr(y), B, T, self.fft_size, self.hop, self.win_length,
             ptr(self.window(x.device)), 1e-7, ptr(sums))

    def backward(self, x, y, sums, w_sc, w_mag, dx):
        B, T = x.shape
        call("artic_stft_loss_bwd", ptr(x), ptr(y), B, T, self.fft_size, self.hop, self.win_length,
             ptr(self.window(x.device)), 1e-7, ptr(sums), float(w_sc), float(w_mag), ptr(dx))


class _MRSTFTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, mod):
        x2, y2 = _as_2d(x), _as_2d(y)
        R = len(mod.resolutions)
        sums = torch.zeros((R, 3), dtype=torch.float32, device=x2.device)
        # the resolutions are independent: one stream each (a 2048-point frame block occupies an SM ~4x longer than a
        # 512-point one; side by side the three launches fill the machine instead of queueing behind each other's tails)
        from ..engine import fork_join
        fork_join([lambda r=r, res=res: res.forward(x2, y2, sums[r]) for r, res in enumerate(mod.resolutions)])
        numel = torch.tensor([res.numel(*x2.shape) for res in mod.resolutions], dtype=torch.float32, device=x2.device)
        sc = (sums[:, 0].sqrt() / sums[:, 1].sqrt()).mean()      # stft_loss.py:61, :166
        mag = (sums[:, 2] / numel).mean()                        # stft_loss.py:82, :167
        ctx.save_for_backward(x2, y2, sums)
        ctx.mod, ctx.shape = mod, x.shape
        return sc, mag

    @staticmethod
    def backward(ctx, g_sc, g_mag):
        x2, y2, sums = ctx.saved_tensors
        R = len(ctx.mod.resolutions)
        dx = torch.zeros_like(x2)
        # the kernel takes host scalars for the two upstream gradients
        w_sc, w_mag = float(g_sc) / R, float(g_mag) / R
        from ..engine import fork_join
        fork_join([lambda r=r, res=res: res.backward(x2, y2, sums[r], w_sc, w_mag, dx)
                   for r, res in enumerate(ctx.mod.resolutions)])          # dx is accumulated atomically
        return dx.view(ctx.shape), None, None


class MultiResolutionSTFTLoss(torch.nn.Module):
    """Drop-in for reference losses/stft_loss.py:121-170."""

    def __init__(self, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
                 window="hann_window"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.resolutions = [STFTResolution(f, h, w, window) for f, h, w in zip(fft_sizes, hop_sizes, win_lengths)]

    def forward(self, x, y):
        """x predicted, y ground truth, (B, T) or (B, C, T) -> (sc_loss, mag_loss)."""
        return _MRSTFTFn.apply(x, y, self)


class MelSpectrogramLoss(torch.nn.Module):
    """Drop-in for reference losses/mel_loss.py:114-166 (center=True, onesided=True,
    normalized=False only — the values the reference configs use)."""

    def __init__(self, fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80,
                 fmin=80, fmax=7600, center=True, normalized=False, onesided=True, eps=1e-10, log_base=10.0):
        super().__init__()
        if not (center and onesided and not normalized):
            raise NotImplementedError("only center=True, onesided=True, normalized=False is implemented")
        if window is not None and not hasattr(torch, f"{window}_window"):
            raise ValueError(f"{window} window is not implemented")
        self.fft_size, self.hop = fft_size, hop_size
        self.win_length = fft_size if win_length is None else win_length
        self.eps, self.num_mels = eps, num_mels
        if log_base is None:
            self.log_scale = 1.0
        elif log_base == 2.0:
            self.log_scale = 1.0 / math.log(2.0)
        elif log_base == 10.0:
            self.log_scale = 1.0 / math.log(10.0)
        else:
            raise ValueError(f"log_base: {log_base} is not supported.")
        fmin = 0 if fmin is None else fmin
        fmax = fs / 2 if fmax is None else fmax
        w = getattr(torch, f"{window}_window")(self.win_length) if window is not None else torch.ones(self.win_length)
        self.register_buffer("window", w.float(), persistent=False)
        melmat = slaney_mel_basis(fs, fft_size, num_mels, fmin, fmax)            # (mels, bins)
        self.register_buffer("melmat", torch.from_numpy(np.ascontiguousarray(melmat.T)).float())  # (bins, mels)
        # non-zero ranges of the (triangular, ~97 % zero) filters: per filter the bins, per bin the filters
        nz = np.asarray(melmat) != 0
        rng = []
        for rows in (nz, nz.T):
            for r in rows:
                idx = np.nonzero(r)[0]
                rng += [int(idx[0]), int(idx[-1]) + 1] if idx.size else [0, 0]
        self.register_buffer("mel_ranges", torch.tensor(rng, dtype=torch.int32), persistent=False)

    def numel(self, B, T):
        return B * self.num_mels * (1 + T // self.hop)

    def accumulate(self, x, y, scale, slot):
        B, T = x.shape
        call("artic_mel_loss_fwd", ptr(x), ptr(y), B, T, self.fft_size, self.hop, self.win_length,
             ptr(self.window), ptr(self.melmat), self.num_mels, self.eps, self.log_scale, float(scale), ptr(slot))

    def backward_into(self, x, y, scale, dx):
        B, T = x.shape
        call("artic_mel_loss_bwd", ptr(x), ptr(y), B, T, self.fft_size, self.hop, self.win_length,
             ptr(self.window), ptr(self.melmat), self.num_mels, self.eps, self.log_scale, float(scale), ptr(dx))

    def loss_and_grad(self, x, y, loss_scale, slot, grad_scale, dx):
        """slot[0] += loss_scale * sum |..| and dx += grad_scale * d(sum |..|)/dx in one launch."""
        B, T = x.shape
        call("artic_mel_loss_fwd_bwd", ptr(x), ptr(y), B, T, self.fft_size, self.hop, self.win_length,
             ptr(self.window), ptr(self.melmat), ptr(self.mel_ranges), self.num_mels, self.eps, self.log_scale,
             float(loss_scale), ptr(slot), float(grad_scale), ptr(dx))

    def forward(self, y_hat, y):
        return _MelFn.apply(y_hat, y, self)


class _MelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, mod):
        x2, y2 = _as_2d(x), _as_2d(y)
        if mod.window.device != x2.device:
            mod.to(x2.device)
        slot = torch.zeros(1, dtype=torch.float32, device=x2.device)
        mod.accumulate(x2, y2, 1.0 / mod.numel(*x2.shape), slot)
        ctx.save_for_backward(x2, y2)
        ctx.mod, ctx.shape = mod, x.shape
        return slot[0]

    @staticmethod
    def backward(ctx, g):
        x2, y2 = ctx.saved_tensors
        dx = torch.zeros_like(x2)
        ctx.mod.backward_into(x2, y2, float(g) / ctx.mod.numel(*x2.shape), dx)
        return dx.view(ctx.shape), None, None
