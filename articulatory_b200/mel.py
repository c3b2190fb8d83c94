"""Slaney-scale mel filterbank used by MelSpectrogramLoss.

The reference obtains it from ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)``
(losses/mel_loss.py:53-59; librosa 0.8.1 defaults htk=False, norm='slaney').  librosa is
not a dependency of this package; the filterbank is a one-off host-side table (float64
arithmetic, float32 result) built from the published formula.
"""
import numpy as np


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    safe = np.maximum(f, 1000.0)
    return np.where(f >= 1000.0, 15.0 + 27.0 * np.log(safe / 1000.0) / np.log(6.4), lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((m - 15.0) * np.log(6.4) / 27.0), m * 200.0 / 3.0)


def slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """(n_mels, n_fft//2 + 1) float32 triangular filters with Slaney area normalisation."""
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    out = np.zeros((n_mels, bins.size), dtype=np.float32)
    for i in range(n_mels):
        lo, mid, hi = edges[i], edges[i + 1], edges[i + 2]
        rise = (bins - lo) / (mid - lo)
        fall = (hi - bins) / (hi - mid)
        out[i] = np.maximum(0.0, np.minimum(rise, fall))
    out *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return out
