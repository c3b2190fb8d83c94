"""Import shim for the UNMODIFIED reference package at /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the
restatement in ``oracle/torch_oracle.py`` and (b) generate the golden fixtures
under ``tests/golden/``.  /root/reference does not exist on the GPU box, so
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file.

The reference imports a handful of packages that are absent in this image
(h5py, librosa, soundfile, tensorboardX, kaldiio, matplotlib, resampy, tkinter)
and one removed scipy alias (scipy.signal.kaiser, layers/pqmf.py:12).  None of
them is on the hot path except ``librosa.filters.mel`` (losses/mel_loss.py:53),
which is stubbed with ``oracle.mel_basis.slaney_mel_basis`` (an independent
float64 restatement of the published librosa 0.8.1 algorithm).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("ARTIC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "articulatory"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Install the stubs and put the reference on sys.path. Idempotent."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    import scipy.signal
    import scipy.signal.windows

    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser

    from oracle.mel_basis import slaney_mel_basis

    def _mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
        return slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)

    for name in ("h5py", "soundfile", "resampy", "kaldiio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    if "librosa" not in sys.modules:
        lib = _stub("librosa")
        lib.filters = _stub("librosa.filters", mel=_mel)
        lib.util = _stub("librosa.util")
    if "tkinter" not in sys.modules or not hasattr(sys.modules["tkinter"], "X"):
        try:
            import tkinter  # noqa: F401
        except Exception:
            _stub("tkinter", X=None)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
    try:
        import tensorboardX  # noqa: F401
    except Exception:
        class SummaryWriter:  # minimal stand-in (bin/train.py:110)
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

        _stub("tensorboardX", SummaryWriter=SummaryWriter)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
