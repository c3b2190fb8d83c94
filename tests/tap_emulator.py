"""Numpy emulator of the artic_tapconv / artic_tapconv_wgrad / artic_weight_prep CONTRACT
(include/artic.h) — test infrastructure for the host-side index algebra (no GPU)."""
import numpy as np


def prep_weight(w, spec, direction):
    """torch-layout weight (numpy) -> prepared [K][G][A][B] following ConvSpec.prep_strides."""
    A, B, sk, sg, sa, sb = spec.prep_strides(direction)
    flat = np.ascontiguousarray(w).reshape(-1)
    K, G = spec.k, spec.groups
    out = np.zeros((K, G, A, B), dtype=w.dtype)
    for k in range(K):
        for g in range(G):
            idx = k * sk + g * sg + (np.arange(A)[:, None] * sa + np.arange(B)[None, :] * sb)
            out[k, g] = flat[idx]
    return out


def unprep_weight(dwp, spec, direction="fwd"):
    """Inverse permutation: prepared-layout gradient -> torch layout."""
    A, B, sk, sg, sa, sb = spec.prep_strides(direction)
    flat = np.zeros(int(np.prod(spec.weight_shape())), dtype=dwp.dtype)
    for k in range(spec.k):
        for g in range(spec.groups):
            idx = k * sk + g * sg + (np.arange(A)[:, None] * sa + np.arange(B)[None, :] * sb)
            flat[idx] = dwp[k, g]
    return flat.reshape(spec.weight_shape())


def tapconv(X, Wp, launches, ly, bias=None):
    """X: (N, Lx, G*A) channels-last; Wp: [K][G][A][B]; returns Y (N, ly, G*B)."""
    N, Lx, _ = X.shape
    K, G, A, B = Wp.shape
    Y = np.zeros((N, ly, G * B), dtype=np.float64)
    written = np.zeros(ly, dtype=bool)
    for L in launches:
        for qi in range(L.nq):
            q = L.q0 + qi
            row = q * L.so + L.ro
            if row < 0 or row >= ly:
                continue
            assert not written[row], "output row written twice"
            written[row] = True
            for off, wi in zip(L.off, L.widx):
                pos = q * L.si + off
                if pos < 0 or pos >= Lx:
                    continue
                for g in range(G):
                    Y[:, row, g * B:(g + 1) * B] += X[:, pos, g * A:(g + 1) * A] @ Wp[wi, g]
    assert written.all(), "some output rows never written"
    if bias is not None:
        Y += bias[None, None, :]
    return Y


def tapwgrad(X, dY, L, K, G, A, B):
    """dW[K][G][A][B] per the artic_tapconv_wgrad contract."""
    N, Lx, _ = X.shape
    Ly = dY.shape[1]
    dW = np.zeros((K, G, A, B), dtype=np.float64)
    for qi in range(L.nq):
        q = L.q0 + qi
        for off, yoff, wi in zip(L.off, L.yoff, L.widx):
            xp, yp = q * L.si + off, q * L.so + yoff
            if xp < 0 or xp >= Lx or yp < 0 or yp >= Ly:
                continue
            for g in range(G):
                dW[wi, g] += X[:, xp, g * A:(g + 1) * A].T @ dY[:, yp, g * B:(g + 1) * B]
    return dW
