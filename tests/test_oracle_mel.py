"""librosa.filters.mel restatement (oracle/mel_basis.py) — PARITY UNPINNED boundary:
cross-check against torchaudio's independent implementation + properties."""
import numpy as np
import pytest

from oracle.mel_basis import hz_to_mel, mel_to_hz, slaney_mel_basis


def test_shape_and_empty_filters():
    m = slaney_mel_basis(16000, 1024, 80, 0, 11025)
    assert m.shape == (80, 513) and m.dtype == np.float32
    assert (m >= 0).all()
    # fmax 11025 > Nyquist 8000: the top filters are empty
    assert int((m.sum(1) == 0).sum()) == 6
    # triangular: each bin is covered by at most two filters
    assert int((m > 0).sum(0).max()) <= 2


def test_mel_scale_roundtrip():
    f = np.array([0.0, 200.0, 999.0, 1000.0, 4000.0, 11025.0])
    assert np.allclose(mel_to_hz(hz_to_mel(f)), f, rtol=1e-12, atol=1e-9)
    assert abs(float(hz_to_mel(1000.0)) - 15.0) < 1e-12


def test_against_torchaudio():
    ta = pytest.importorskip("torchaudio")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t = ta.functional.melscale_fbanks(513, 0.0, 11025.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(slaney_mel_basis(16000, 1024, 80, 0, 11025) - t).max() < 2e-7
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t = ta.functional.melscale_fbanks(257, 80.0, 7600.0, 40, 22050, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(slaney_mel_basis(22050, 512, 40, 80, 7600) - t).max() < 2e-7
