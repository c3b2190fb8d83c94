"""Slaney mel filterbank — restatement of librosa.filters.mel (TEST INFRASTRUCTURE).

The reference calls ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)``
(reference articulatory/losses/mel_loss.py:53-59; pinned ``librosa==0.8.1`` in
requirements.txt:35).  librosa is a third-party dependency that is NOT vendored in
/root/reference and is not installed in this image, so its published algorithm
(librosa 0.8.1 ``filters.mel`` with the defaults ``htk=False, norm='slaney'``)
is restated here in numpy:

  * Slaney mel scale: linear below 1 kHz with f_sp = 200/3 Hz per mel, then
    logarithmic with step ln(6.4)/27 per mel,
  * n_mels + 2 band edges equally spaced in mel between fmin and fmax,
  * triangular weights max(0, min(lower_ramp, upper_ramp)) on the rFFT bin
    centre frequencies, built in float64 and stored as float32,
  * Slaney area normalisation 2 / (f[i+2] - f[i]).

PARITY UNPINNED at this boundary: the reference holds no test or golden vector for
the filterbank.  It is cross-checked in tests/test_oracle_mel.py against
``torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")`` (an
independent implementation of the same published formula) and by property tests.
"""
import numpy as np

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    mel = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    safe = np.where(log_t, f, _MIN_LOG_HZ)
    return np.where(log_t, _MIN_LOG_MEL + np.log(safe / _MIN_LOG_HZ) / _LOGSTEP, mel)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), f)


def slaney_mel_basis(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """Returns the (n_mels, 1 + n_fft//2) float32 filterbank."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights
