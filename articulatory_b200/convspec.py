"""Host-side description of one conv-like layer and its mapping onto the tap-gather GEMM
(include/artic.h: artic_tapconv / artic_tapconv_wgrad).

Pure Python / integer arithmetic: no CUDA needed to BUILD a plan, so the index algebra is
unit-tested on CPU against torch's conv semantics (tests/test_convspec.py) with a small
numpy emulator of the kernel contract.

Layer kinds (reference call sites):
  conv   — torch.nn.Conv1d / Conv2d(k,1)   weight (Co, Ci/g, k[,1])  models/hifigan.py:108-172,365-388,549-615
  convT  — torch.nn.ConvTranspose1d        weight (Ci, Co, k)        models/hifigan.py:124-131
  linear — torch.nn.Linear                 weight (out, in)          layers/pytorch_layers.py:437-449
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

FAR = -(1 << 28)  # a tap offset that is always out of range (reads zero)


@dataclass
class TapLaunch:
    """One artic_tapconv launch: out[row = q*so + ro] = sum_t X[q*si + off[t]] @ W[widx[t]]."""
    q0: int
    nq: int
    si: int
    so: int
    ro: int
    off: List[int]
    widx: List[int]


@dataclass
class WgradLaunch:
    q0: int
    nq: int
    si: int
    so: int
    off: List[int]
    yoff: List[int]
    widx: List[int]


@dataclass
class ConvSpec:
    kind: str                 # 'conv' | 'convT' | 'linear'
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    dilation: int = 1
    padding: int = 0
    output_padding: int = 0
    groups: int = 1
    weight_norm: bool = False
    name: str = ""
    trim_right: int = 0       # output samples dropped at the end (causal convs, layers/causal_conv.py:42,66)
    # derived
    cig: int = field(init=False)
    cog: int = field(init=False)

    def __post_init__(self):
        assert self.kind in ("conv", "convT", "linear")
        assert self.cin % self.groups == 0 and self.cout % self.groups == 0
        if self.kind != "conv":
            assert self.groups == 1 and self.dilation == 1
        if self.kind == "linear":
            assert self.k == 1 and self.stride == 1 and self.padding == 0
        assert self.k <= 48
        self.cig = self.cin // self.groups
        self.cog = self.cout // self.groups

    # ---- geometry ----------------------------------------------------------------
    def out_len(self, lin: int) -> int:
        if self.kind == "convT":
            return (lin - 1) * self.stride - 2 * self.padding + self.k + self.output_padding - self.trim_right
        return (lin + 2 * self.padding - self.dilation * (self.k - 1) - 1) // self.stride + 1 - self.trim_right

    # ---- torch weight layout -----------------------------------------------------
    def weight_shape(self) -> Tuple[int, ...]:
        if self.kind == "conv":
            return (self.cout, self.cig, self.k)
        if self.kind == "convT":
            return (self.cin, self.cout, self.k)
        return (self.cout, self.cin)

    def wn_rows(self) -> Tuple[int, int]:
        """(rows, row_len) of the torch weight seen as [dim0][rest] (weight-norm dim 0)."""
        shp = self.weight_shape()
        rest = 1
        for s in shp[1:]:
            rest *= s
        return shp[0], rest

    def prep_strides(self, direction: str) -> Tuple[int, int, int, int, int, int]:
        """(A, B, sk, sg, sa, sb) of the prepared weight [K][G][A][B] for 'fwd' (A = in
        channels/group, B = out channels/group) or 'bwd' (A = out, B = in)."""
        K = self.k
        if self.kind == "conv":
            s_ci, s_co, s_g = K, self.cig * K, self.cog * self.cig * K
            sk = 1
        elif self.kind == "convT":
            s_ci, s_co, s_g, sk = self.cout * K, K, 0, 1
        else:
            s_ci, s_co, s_g, sk = 1, self.cin, 0, 0
        if direction == "fwd":
            return self.cig, self.cog, sk, s_g, s_ci, s_co
        return self.cog, self.cig, sk, s_g, s_co, s_ci

    # ---- launches ----------------------------------------------------------------
    def fwd_launches(self, lin: int) -> List[TapLaunch]:
        lout = self.out_len(lin)
        if self.kind in ("conv", "linear"):
            offs = [j * self.dilation - self.padding for j in range(self.k)]
            return [TapLaunch(0, lout, self.stride, 1, 0, offs, list(range(self.k)))]
        return _phase_launches(self.k, self.stride, 1, self.padding, lout)

    def dgrad_launches(self, lin: int) -> List[TapLaunch]:
        """Launches computing dX (len lin) from dY (len out_len(lin)); use the 'bwd' weight."""
        if self.kind in ("conv", "linear"):
            return _phase_launches(self.k, self.stride, self.dilation, self.padding, lin)
        # convT backward-data is a plain strided conv over dY
        offs = [j - self.padding for j in range(self.k)]
        return [TapLaunch(0, lin, self.stride, 1, 0, offs, list(range(self.k)))]

    def wgrad_launch(self, lin: int) -> WgradLaunch:
        """dW in the 'fwd' prepared layout [K][G][cig][cog]; X = layer input, dY = output grad."""
        lout = self.out_len(lin)
        K, s = self.k, self.stride
        if self.kind in ("conv", "linear"):
            return WgradLaunch(0, lout, s, 1, [j * self.dilation - self.padding for j in range(K)],
                               [0] * K, list(range(K)))
        return WgradLaunch(0, lin + (K - 1) // s, 1, s, [-(j // s) for j in range(K)],
                           [(j % s) - self.padding for j in range(K)], list(range(K)))


def _phase_launches(k: int, s: int, d: int, pad: int, lrows: int) -> List[TapLaunch]:
    """Transposed-direction gather: out[u] = sum_j in[(u + pad - j*d)/s] W_j (when divisible),
    u in [0, lrows).  One launch per residue r = (u + pad) mod s:  u = s*q + r - pad."""
    out = []
    for r in range(s):
        taps = [j for j in range(k) if (j * d) % s == r]
        q_lo = -((r - pad) // s)            # ceil((pad - r)/s)
        q_hi = (lrows - 1 + pad - r) // s   # floor
        if q_hi < q_lo:
            continue
        if taps:
            offs = [-((j * d - r) // s) for j in taps]
            widx = taps
        else:                               # no tap lands on this phase: write zeros (+ epilogue)
            offs, widx = [FAR], [0]
        out.append(TapLaunch(q_lo, q_hi - q_lo + 1, 1, s, r - pad, offs, widx))
    return out
