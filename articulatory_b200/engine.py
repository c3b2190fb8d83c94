"""Host-side execution engine: explicit forward / backward schedules of the HiFi-GAN
generator and the multi-scale / multi-period discriminator over the C-ABI kernels.

There is no autograd inside: every activation that the backward needs is kept on a
tape by the forward, and the backward is a hand-written schedule of tap-gather GEMM
dgrad / wgrad launches with the LeakyReLU masks, residual adds, MRF 1/3 scaling and
feature-matching gradients fused into the GEMM epilogues.  All work is enqueued on the
current CUDA stream with no host synchronisation, so a whole train step can be captured
in one CUDA graph (trainer.py).

Layouts: activations are channels-last (N, L, C); the period discriminators keep the
reference's (B, H, p) element order with C innermost, i.e. (B, H, p, C), addressed as
N = B*p interleaved sequences (artic_seq_t.n_inner = p) — so the reference's
view(b, c, t//p, p) (models/hifigan.py:417) is a zero-copy reinterpretation.
"""
import contextlib
import os as _os
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, BF16, F32, TORCH_DTYPE, call, ptr
from .convspec import ConvSpec


# --------------------------------------------------------------------------- #
# tensors                                                                     #
# --------------------------------------------------------------------------- #
class SeqT:
    """Channels-last sequence batch in a torch tensor (see include/artic.h artic_seq_t).

    bf16x3 mode: an fp32 batch may carry a SPLIT COPY ``sp`` — a bf16 tensor of shape (2,) + t.shape holding
    hi = bf16(t) and lo = bf16(t - hi), the operand format of the error-compensated tensor-core contraction.  It is
    written by the producing kernel (artic_tapconv_t.Y_sp) or on demand (``ensure_split``); views of the same
    storage (``flat_period``) share it through ``_h``."""
    __slots__ = ("t", "N", "L", "C", "n_inner", "s_outer", "s_inner", "s_row", "code", "_h")

    def __init__(self, t, N, L, C, n_inner=1, s_outer=None, s_inner=0, s_row=None, _h=None):
        self.t, self.N, self.L, self.C, self.n_inner = t, N, L, C, n_inner
        self.s_row = C if s_row is None else s_row
        self.s_outer = L * C if s_outer is None else s_outer
        self.s_inner = s_inner
        self.code = _lib.DTYPE_CODE[t.dtype]
        self._h = [None] if _h is None else _h

    @property
    def sp(self):
        return self._h[0]

    def split_out(self):
        """The split-copy planes of this batch, allocated on first use (for a producer to fill)."""
        if self._h[0] is None:
            assert self.code == F32 and self.t.is_contiguous()
            self._h[0] = torch.empty((2,) + tuple(self.t.shape), dtype=torch.bfloat16, device=self.t.device)
        return self._h[0]

    def ensure_split(self):
        """Split copy of the CURRENT contents (one elementwise launch on the current stream unless a producer
        already wrote it).  Call it on the stream that produced ``t``, before any fork that consumes it."""
        if self._h[0] is None:
            sp = self.split_out()
            call("artic_split", ptr(self.t), ptr(sp), sp.stride(0), self.t.numel())
        return self._h[0]

    @staticmethod
    def empty(N, L, C, code, device, zero=False):
        f = torch.zeros if zero else torch.empty
        return SeqT(f((N, L, C), dtype=TORCH_DTYPE[code], device=device), N, L, C)

    @staticmethod
    def period(B, H, p, C, code, device):
        t = torch.empty((B, H, p, C), dtype=TORCH_DTYPE[code], device=device)
        return SeqT(t, B * p, H, C, n_inner=p, s_outer=H * p * C, s_inner=C, s_row=p * C)

    def like(self, code=None, C=None, L=None):
        """New storage with the same batch structure (optionally other dtype / width / length)."""
        code = self.code if code is None else code
        C = self.C if C is None else C
        L = self.L if L is None else L
        if self.n_inner == 1:
            return SeqT.empty(self.N, L, C, code, self.t.device)
        return SeqT.period(self.N // self.n_inner, L, self.n_inner, C, code, self.t.device)

    def seq(self):
        return _lib.Seq(self.s_outer, self.s_inner, self.s_row, self.n_inner, self.L)

    def numel(self):
        return self.N * self.L * self.C


def _fill_taps(dst_off, dst_idx, off, widx):
    n = len(off)
    for i in range(n):
        dst_off[i] = off[i]
        dst_idx[i] = widx[i]
    return n


def tapconv_params(launches, X: SeqT, W, G, Cig, Cog, Y: Optional[SeqT] = None, Y2: Optional[SeqT] = None,
                   bias=None, res_pre: Optional[SeqT] = None, mask: Optional[SeqT] = None,
                   res: Optional[SeqT] = None, res2: Optional[SeqT] = None, alpha=1.0, mask_slope=1.0,
                   act=ACT_NONE, act_slope=0.0, Wt=None, Wt_sp=None, sp_y=False, sp_y2=False):
    """The artic_tapconv_t parameter blocks of one layer direction (one per launch phase).  ``Wt`` is the
    same weight in the transposed prepared layout [K][G][Cog][Cig] (enables the tcgen05 kernel).
    bf16x3 mode: ``Wt_sp`` = (split copy of Wt: bf16 hi plane, element distance to the lo plane); X then travels
    with its split copy, and ``sp_y`` / ``sp_y2`` ask for split copies of Y / Y2 (outputs that already own one are
    always kept in sync)."""
    yref = Y if Y is not None else Y2
    assert yref is not None and X.C == G * Cig and yref.C == G * Cog and X.N == yref.N
    for o in (Y, Y2, res_pre, mask, res, res2):
        assert o is None or (o.code == yref.code and o.s_row == yref.s_row and o.s_outer == yref.s_outer
                             and o.L == yref.L and o.n_inner == yref.n_inner), "epilogue tensors must share Y's layout"
    assert W.dtype == X.t.dtype
    out = []
    for L in launches:
        p = _lib.TapConv()
        p.X, p.W, p.bias = ptr(X.t), ptr(W), ptr(bias)
        p.res_pre = ptr(res_pre.t) if res_pre is not None else None
        p.mask = ptr(mask.t) if mask is not None else None
        p.res = ptr(res.t) if res is not None else None
        p.res2 = ptr(res2.t) if res2 is not None else None
        p.Y = ptr(Y.t) if Y is not None else None
        p.Y2 = ptr(Y2.t) if Y2 is not None else None
        p.x, p.y = X.seq(), yref.seq()
        p.N, p.G, p.Cig, p.Cog = X.N, G, Cig, Cog
        p.q0, p.nq, p.si, p.so, p.ro = L.q0, L.nq, L.si, L.so, L.ro
        p.ntaps = _fill_taps(p.off, p.widx, L.off, L.widx)
        p.alpha, p.mask_slope, p.act_slope, p.act = alpha, mask_slope, act_slope, act
        p.dtype, p.out_dtype = X.code, yref.code
        if Wt is not None and Wt.dtype == W.dtype:
            p.Wt, p.Wt_taps = ptr(Wt), Wt.shape[0]
        if Wt_sp is not None and X.code == F32 and yref.code == F32:
            xs = X.ensure_split()
            p.X_sp, p.x_plane = ptr(xs), xs.stride(0)
            p.Wt_sp, p.w_plane = ptr(Wt_sp[0]), Wt_sp[1]
        for o, want, field in ((Y, sp_y, "Y_sp"), (Y2, sp_y2, "Y2_sp")):
            if o is not None and o.code == F32 and (want or o.sp is not None):
                so = o.split_out()
                assert p.y_plane in (0, so.stride(0)), "Y and Y2 split copies must share their plane distance"
                setattr(p, field, ptr(so))
                p.y_plane = so.stride(0)
        out.append(p)
    return out


def launch_tapconvs(params):
    """Enqueue a list of INDEPENDENT artic_tapconv_t problems with as few launches as possible."""
    for i in range(0, len(params), 64):
        chunk = params[i:i + 64]
        arr = (_lib.TapConv * len(chunk))(*chunk)
        call("artic_tapconv_multi", arr, len(chunk))


def tapconv(launches, X: SeqT, W, G, Cig, Cog, **kw):
    """Issue the launch phases of one layer direction (one multi-problem call)."""
    launch_tapconvs(tapconv_params(launches, X, W, G, Cig, Cog, **kw))


# --------------------------------------------------------------------------- #
# one conv-like layer                                                         #
# --------------------------------------------------------------------------- #

#: ARTIC_GROUP=1: the three MRF blocks of a generator stage advance in lock-step through multi-problem
#: launches instead of on three concurrent streams (measured on B200: 17.8 vs 17.7 ms per step, i.e. no
#: better — the step is bound by SM time, not by launch count — so the stream schedule stays the default)
_GROUPED = _os.environ.get("ARTIC_GROUP", "0") == "1"
_SIDE_STREAMS: Dict[int, List["torch.cuda.Stream"]] = {}
_FORK_DEPTH: Dict[int, int] = {}


#: Stream priorities: everything on the critical path runs on HIGH-priority streams (the capture stream
#: of trainer.TrainStep, fork_join side streams, weight-gradient queues); "background" branches — the
#: discriminator's weight re-materialisation under the generator forward, the spectral losses under the
#: discriminator forward — run on LOW-priority streams, so that their thread blocks only take SMs the
#: critical-path kernels do not want (measured with tools/kernel_trace.py: without this the 1184 blocks of
#: the weight prep / the 1600 frame-FFT blocks of the mel loss fill every SM and the chain they should
#: overlap with starts 300-500 us late).
PRIO_HIGH, PRIO_LOW = -1, 0
_BG_STREAMS: Dict[int, List["torch.cuda.Stream"]] = {}
_BG_DEPTH: Dict[int, int] = {}
_IN_BACKGROUND = [False]


def critical_stream(device=None):
    """A new high-priority stream (for graph capture / the eager step)."""
    return torch.cuda.Stream(device=device, priority=PRIO_HIGH)


def fork_join(branches, device=None, background=()):
    """Run independent branches concurrently: branch 0 on the current stream, the others on side
    streams forked from it, all joined back before returning (capturable in a CUDA graph: the
    fork / join become graph edges).  Branch indices listed in ``background`` (and every fork_join
    nested inside them) use low-priority streams.  The layers of this network are small enough that a single
    kernel leaves SMs idle and pays its launch / pipeline-fill latency serially; the three MRF
    blocks of a generator stage and the eight sub-discriminators are independent, so their kernels
    overlap.  ARTIC_STREAMS=0 runs the branches one after the other."""
    os = _os
    if len(branches) == 1 or os.environ.get("ARTIC_STREAMS", "1") == "0":
        return [b() for b in branches]
    main = torch.cuda.current_stream()
    dev = main.device_index
    use_prio = os.environ.get("ARTIC_PRIORITIES", "1") != "0"
    pool = _SIDE_STREAMS.setdefault(dev, [])
    bg_pool = _BG_STREAMS.setdefault(dev, [])
    base, bg_base = _FORK_DEPTH.get(dev, 0), _BG_DEPTH.get(dev, 0)   # nested fork_join calls take fresh side streams
    side = []
    n_hi = n_bg = 0
    for i in range(1, len(branches)):
        if use_prio and (_IN_BACKGROUND[0] or i in background):
            if len(bg_pool) <= bg_base + n_bg:
                bg_pool.append(torch.cuda.Stream(device=dev, priority=PRIO_LOW))
            side.append(bg_pool[bg_base + n_bg])
            n_bg += 1
        else:
            if len(pool) <= base + n_hi:
                pool.append(torch.cuda.Stream(device=dev, priority=PRIO_HIGH if use_prio else 0))
            side.append(pool[base + n_hi])
            n_hi += 1
    _FORK_DEPTH[dev], _BG_DEPTH[dev] = base + n_hi, bg_base + n_bg
    was_bg = _IN_BACKGROUND[0]
    try:
        for st in side:
            st.wait_stream(main)
        out = [None] * len(branches)
        for i in range(1, len(branches)):
            _IN_BACKGROUND[0] = was_bg or i in background
            with torch.cuda.stream(side[i - 1]):
                out[i] = branches[i]()
        _IN_BACKGROUND[0] = was_bg or 0 in background
        out[0] = branches[0]()
        for st in side:
            main.wait_stream(st)
    finally:
        _FORK_DEPTH[dev], _BG_DEPTH[dev] = base, bg_base
        _IN_BACKGROUND[0] = was_bg
    return out


_QUEUE_STREAMS: Dict[int, List["torch.cuda.Stream"]] = {}
_QUEUE_NEXT: Dict[int, int] = {}


class SideQueue:
    """A side stream for work that hangs OFF the critical path of the current stream: the weight
    gradients of a backward chain only feed the final weight-norm backward, so they are queued here
    and the data-gradient chain does not wait for them.  Every submission first waits for the work
    already enqueued on the submitting stream (its inputs); ``join`` makes the current stream wait
    for the queue.  Tensors handed to queued kernels must stay referenced until ``join`` (``keep``)."""

    def __init__(self):
        import os
        self.inline = os.environ.get("ARTIC_STREAMS", "1") == "0"
        self.keep = []
        if not self.inline:
            dev = torch.cuda.current_stream().device_index
            use_prio = os.environ.get("ARTIC_PRIORITIES", "1") != "0"
            bg = use_prio and _IN_BACKGROUND[0]        # queues of a background branch stay in the background
            prio = (PRIO_LOW if bg else PRIO_HIGH) if use_prio else 0
            key = (dev, bg)
            pool = _QUEUE_STREAMS.setdefault(key, [torch.cuda.Stream(device=dev, priority=prio) for _ in range(12)])
            i = _QUEUE_NEXT.get(key, 0)
            _QUEUE_NEXT[key] = (i + 1) % len(pool)
            self.s = pool[i]

    def run(self, fn, *tensors):
        if self.inline:
            return fn()
        self.keep.extend(tensors)
        self.s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.s):
            fn()

    def join(self):
        if not self.inline:
            torch.cuda.current_stream().wait_stream(self.s)
        keep, self.keep = self.keep, []
        return keep


def slice_seq(s: SeqT, lo: int, hi: int) -> SeqT:
    """View of batch items [lo, hi) (the batch is the leading tensor dim in every layout)."""
    h = [s.sp[:, lo:hi]] if s.sp is not None else None
    return SeqT(s.t[lo:hi], (hi - lo) * s.n_inner, s.L, s.C, s.n_inner, s.s_outer, s.s_inner, s.s_row, _h=h)


def flat_period(s: SeqT) -> SeqT:
    """(B, H, p, C) period storage seen as B plain sequences of H*p rows (zero-copy).  A stride-1
    Conv2d (k,1) over (H, p) is exactly a Conv1d over that flattened length with dilation p and padding
    pad*p: one dense tile per batch item instead of p short interleaved sequences."""
    if s.n_inner == 1:
        return s
    B = s.N // s.n_inner
    return SeqT(s.t, B, s.L * s.n_inner, s.C, n_inner=1, s_outer=s.s_outer, s_inner=0, s_row=s.s_inner, _h=s._h)


class ConvLayer:
    """A conv / convT / linear layer bound to its torch parameters.

    ``in_code`` is the storage dtype of the layer input (and of the forward weight),
    ``out_code`` the dtype of its output (and of the dY / backward weight)."""

    def __init__(self, spec: ConvSpec, name: str, in_code: int, out_code: int, pad_in: bool = False, x3: bool = False):
        self.spec, self.name, self.in_code, self.out_code = spec, name, in_code, out_code
        # bf16x3 mode: fp32 storage, split-operand tensor-core contraction (artic.h: artic_tapconv_t.X_sp)
        self.x3 = bool(x3) and in_code == F32 and out_code == F32
        tc = self.x3 or (in_code == BF16 and out_code == BF16)     # layer may run on the tcgen05 kernels
        # Narrow grouped convs (< 32 channels per group, e.g. the scale discriminator's 128->256
        # g16 layer with 8 -> 16 channels per group) would starve the tensor-core tiles: in the
        # bf16 mode `mg` groups are merged into one super-group whose weight is block-diagonal
        # (artic_wdesc_t.merge), trading mg x redundant MACs for full MMA tiles.
        mg = 1
        if tc and spec.groups > 1:
            while (spec.cig * mg < 32 or spec.cog * mg < 32) and spec.groups % (2 * mg) == 0:
                mg *= 2
        self.mg = mg
        self.kG, self.kcig, self.kcog = spec.groups // mg, spec.cig * mg, spec.cog * mg   # kernel-facing dims
        # Odd input widths (the generator's 13 + 128 = 141-channel input conv) are zero-padded to a
        # multiple of 32 channels so that the layer runs on the tensor-core kernel.
        if pad_in and tc and spec.groups == 1 and spec.cin >= 32 and spec.cin % 16:
            self.kcig = (spec.cin + 31) // 32 * 32
            if self.kcig > 128:      # the tcgen05 weight-gradient kernel wants 32 / 64 / k * 128 input channels
                self.kcig = (spec.cin + 127) // 128 * 128
        # Transposed convs get their weight gradient as the weight gradient of the equivalent strided conv
        # with X and dY exchanged (dW^T, i.e. the 'bwd' layout): that contraction has unit output stride,
        # which the tcgen05 wgrad kernel requires.
        if self.x3 and (self.kcig % 16 or self.kcog % 32):
            self.x3 = False                      # channel-1 ends / odd widths stay on the fp32 CUDA-core kernels
            tc = False
        self.dw_swapped = spec.kind == "convT" and tc
        self.v = self.g = self.b = None          # torch parameters (fp32, device)
        self.Wf = self.Wb = self.scale = None    # prepared weights
        self.Wf_sp = self.Wb_sp = None           # bf16x3: (hi plane of the split copy, distance to the lo plane)
        self.dWf = None                          # fp32 wgrad accumulator, 'fwd' layout
        self._own = None                         # single-layer WeightSet (tests / stand-alone use)

    # ---- parameters ----------------------------------------------------------
    def bind(self, params: Dict[str, torch.Tensor]):
        """Bind to tensors from a name->tensor mapping with the reference's key names."""
        n = self.name
        if n + ".weight_v" in params:
            self.v, self.g = params[n + ".weight_v"], params[n + ".weight_g"]
        else:
            self.v, self.g = params[n + ".weight"], None
        self.b = params.get(n + ".bias")
        for t in (self.v, self.g, self.b):
            if t is not None:
                _lib.require_cuda(t, f"parameter of {n}")
                assert t.dtype == torch.float32 and t.is_contiguous()
        shape = tuple(self.v.shape)
        want = self.spec.weight_shape()
        assert shape == want or shape == want + (1,), f"{n}: weight shape {shape} != {want}"
        self._own = None

    def param_names(self):
        n = self.name
        names = [n + ".weight_g", n + ".weight_v"] if self.g is not None else [n + ".weight"]
        if self.b is not None:
            names.append(n + ".bias")
        return names

    def _single(self):
        if self._own is None:
            self._own = WeightSet([self])
        return self._own

    def prep(self, need_bwd=True):
        """(Re)materialise the effective weights from the current parameters."""
        self._single().prep()

    def zero_wgrad(self):
        self._single().zero()

    def finish_grads(self, grads: Dict[str, torch.Tensor]):
        """dW (prepared layout) -> gradients of the torch parameters (weight-norm backward)."""
        self._single().unprep(grads)

    # ---- compute -------------------------------------------------------------
    def forward_params(self, X: SeqT, Y=None, Y2=None, spec=None, **epi):
        s = spec or self.spec
        if self.x3:
            epi.setdefault("sp_y2", True)        # the activated output feeds the next (tensor-core) layer
        return tapconv_params(s.fwd_launches(X.L), X, self.Wf, self.kG, self.kcig, self.kcog, Y=Y, Y2=Y2, bias=self.b,
                              Wt=self.Wb, Wt_sp=self.Wb_sp, **epi)

    def forward(self, X: SeqT, Y=None, Y2=None, **epi):
        launch_tapconvs(self.forward_params(X, Y=Y, Y2=Y2, **epi))

    def dgrad_params(self, dY: SeqT, dX: Optional[SeqT] = None, dX2: Optional[SeqT] = None, spec=None, **epi):
        s = spec or self.spec
        lin = (dX if dX is not None else dX2).L
        if self.x3:
            epi.setdefault("sp_y", True)         # data gradients feed the next dgrad and a weight gradient
        return tapconv_params(s.dgrad_launches(lin), dY, self.Wb, self.kG, self.kcog, self.kcig, Y=dX, Y2=dX2,
                              Wt=self.Wf, Wt_sp=self.Wf_sp, **epi)

    def dgrad(self, dY: SeqT, dX: Optional[SeqT] = None, dX2: Optional[SeqT] = None, **epi):
        launch_tapconvs(self.dgrad_params(dY, dX=dX, dX2=dX2, **epi))

    def wgrad(self, X: SeqT, dY: SeqT, grads: Dict[str, torch.Tensor], spec=None):
        """Accumulate dW (prepared layout) and the bias gradient (into grads[name.bias])."""
        s = spec or self.spec
        p = _lib.TapWgrad()
        if self.dw_swapped:
            # dW^T[j][co][ci] = sum_q dY[q*s + j - pad][co] * X[q][ci]
            p.X, p.dY, p.dW = ptr(dY.t), ptr(X.t), ptr(self.dWf)
            p.x, p.y = dY.seq(), X.seq()
            p.N, p.G, p.Cig, p.Cog = X.N, self.kG, self.kcog, self.kcig
            p.q0, p.nq, p.si, p.so = 0, X.L, s.stride, 1
            p.ntaps = s.k
            for j in range(s.k):
                p.off[j], p.yoff[j], p.widx[j] = j - s.padding, 0, j
            p.dtype, p.y_dtype = dY.code, X.code
        else:
            L = s.wgrad_launch(X.L)
            p.X, p.dY, p.dW = ptr(X.t), ptr(dY.t), ptr(self.dWf)
            p.x, p.y = X.seq(), dY.seq()
            p.N, p.G, p.Cig, p.Cog = X.N, self.kG, self.kcig, self.kcog
            p.q0, p.nq, p.si, p.so = L.q0, L.nq, L.si, L.so
            p.ntaps = len(L.off)
            for i in range(p.ntaps):
                p.off[i], p.yoff[i], p.widx[i] = L.off[i], L.yoff[i], L.widx[i]
            p.dtype, p.y_dtype = X.code, dY.code
        if self.x3 and self.kcig % 32 == 0 and self.kcog % 32 == 0:
            xs, ys = X.ensure_split(), dY.ensure_split()
            if self.dw_swapped:
                xs, ys = ys, xs
            p.X_sp, p.x_plane, p.dY_sp, p.y_plane = ptr(xs), xs.stride(0), ptr(ys), ys.stride(0)
        fuse_bias = self.b is not None and not self.dw_swapped and p.so == 1 and p.q0 + p.yoff[0] == 0 and p.nq == dY.L
        if fuse_bias:       # column sums of dY from one more tensor-core accumulator (or artic_colsum inside the call)
            p.dbias = ptr(grads[self.name + ".bias"])
        call("artic_tapconv_wgrad", p)
        if self.b is not None and not fuse_bias:
            call("artic_colsum", ptr(dY.t), dY.seq(), dY.N, dY.C, dY.code, ptr(grads[self.name + ".bias"]))

#: a discriminator input row with its AvgPool pyramid must fit the one-launch preamble kernel's shared memory
_DISC_PREP_MAX_BYTES = 200 * 1024
_WEIGHTS_GENERIC = _os.environ.get("ARTIC_WEIGHTS_GENERIC", "0") == "1"


class WeightSet:
    """The prepared weights and weight-gradient accumulators of a set of layers, handled by the
    batched kernels (artic_weights_prep / artic_weights_unprep: one launch per pass for the
    whole network instead of three per layer)."""

    def __init__(self, layers: List[ConvLayer]):
        self.layers = list(layers)
        dev = self.layers[0].v.device
        self.dev = dev
        self.any_norm = int(any(l.g is not None for l in self.layers))
        sizes = []
        # bf16x3 layers: prepared fp32 weights in ONE flat buffer, so that one elementwise launch refreshes the
        # split copies (bf16 hi / lo planes, `x3_total` elements apart) of every layer after a weight update
        def up8(v):
            return (v + 7) // 8 * 8

        n3 = sum(2 * up8(l.spec.k * l.kG * l.kcig * l.kcog) for l in self.layers if l.x3)
        self.x3_total = n3
        self.W3 = torch.zeros(n3, dtype=torch.float32, device=dev) if n3 else None
        self.W3_sp = torch.zeros((2, n3), dtype=torch.bfloat16, device=dev) if n3 else None
        o3 = 0
        for l in self.layers:
            s = l.spec
            shape_f = (s.k, l.kG, l.kcig, l.kcog)
            shape_b = (s.k, l.kG, l.kcog, l.kcig)
            n = s.k * l.kG * l.kcig * l.kcog
            if l.x3:
                ob = o3 + up8(n)
                l.Wf, l.Wb = self.W3[o3:o3 + n].view(shape_f), self.W3[ob:ob + n].view(shape_b)
                l.Wf_sp, l.Wb_sp = (self.W3_sp[0, o3:o3 + n], n3), (self.W3_sp[0, ob:ob + n], n3)
                o3 = ob + up8(n)
            else:
                l.Wf = torch.zeros(shape_f, dtype=TORCH_DTYPE[l.in_code], device=dev)
                l.Wb = torch.zeros(shape_b, dtype=TORCH_DTYPE[l.out_code], device=dev)
            l.scale = torch.empty(2 * s.wn_rows()[0], dtype=torch.float32, device=dev) if l.g is not None else None
            sizes.append(n)
        # one flat fp32 accumulator for all weight gradients: a single memset per backward
        self.dW = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        o = 0
        for l, n in zip(self.layers, sizes):
            s = l.spec
            l.dWf = self.dW[o:o + n].view(*((s.k, l.kG, l.kcog, l.kcig) if l.dw_swapped else (s.k, l.kG, l.kcig, l.kcog)))
            o += n
        self._bufs = [(l.Wf, l.Wb, l.scale, l.dWf, l.Wf_sp, l.Wb_sp) for l in self.layers]
        self._tables = {}

    def _table(self, grads):
        key = None if grads is None else tuple(ptr(grads[n]) for l in self.layers for n in l.param_names()[:1])
        tab = self._tables.get(key)
        if tab is None:
            descs = []
            tiles = tiles2 = 0
            lib = _lib.load()
            # inner tiles per work unit of the row-run kernels: large nets amortise the per-unit index arithmetic over
            # up to 8 tiles, small ones keep one tile per unit so that the grid still fills the machine
            single = sum(lib.artic_wrow_tiles(l.spec.k, l.spec.groups, *l.spec.prep_strides("fwd")[:2],
                                              *(l.spec.prep_strides("fwd")[i] for i in (2, 4, 5)), 1) for l in self.layers)
            chunk = max(1, min(8, single // 2400))
            for l in self.layers:
                s = l.spec
                rows, row_len = s.wn_rows()
                A, B, sk, sg, sa, sb = s.prep_strides("fwd")
                d = _lib.WDesc()
                d.v, d.g, d.scale = ptr(l.v), ptr(l.g), ptr(l.scale)
                d.out_f, d.out_b, d.dWp = ptr(l.Wf), ptr(l.Wb), ptr(l.dWf)
                if grads is not None:
                    n = l.name
                    if l.g is not None:
                        d.dv, d.dg = ptr(grads[n + ".weight_v"]), ptr(grads[n + ".weight_g"])
                    else:
                        d.dv, d.dg = ptr(grads[n + ".weight"]), None
                d.row_len, d.sk, d.sg, d.sa, d.sb = row_len, sk, sg, sa, sb
                d.rows, d.K, d.G, d.A, d.B = rows, s.k, s.groups, A, B
                d.merge, d.a_pad, d.b_pad = l.mg, l.kcig, l.kcog
                d.dtype_f, d.dtype_b = l.in_code, l.out_code
                d.dw_swapped = int(l.dw_swapped)
                d.tile_begin, d.tile2_begin = tiles, tiles2
                # row-run kernels (the default; ARTIC_WEIGHTS_GENERIC=1 forces the generic tile kernels) ...
                n2 = 0 if _WEIGHTS_GENERIC else lib.artic_wrow_tiles(s.k, s.groups, A, B, sk, sa, sb, chunk)
                d.row_chunk = chunk
                if n2 > 0:
                    tiles2 += n2
                else:                                                           # ... or the generic tile kernels
                    tiles += lib.artic_wperm_tiles(s.k, s.groups, A, B)
                descs.append(d)
            self.total_tiles, self.total_tiles2 = tiles, tiles2
            tab = _lib.upload_structs(descs, self.dev)
            if len(self._tables) > 4:       # the autograd path hands in fresh grads every backward
                self._tables = {k: v for k, v in self._tables.items() if k is None}
            self._tables[key] = tab
        return tab

    def rebind(self):
        """Re-attach the buffers to the layers after ConvLayer.bind() (same parameter storage)."""
        for l, (wf, wb, sc, dw, wfs, wbs) in zip(self.layers, self._bufs):
            l.Wf, l.Wb, l.scale, l.dWf, l.Wf_sp, l.Wb_sp = wf, wb, sc, dw, wfs, wbs

    def prep(self):
        tab = self._table(None)
        call("artic_weights_prep", ptr(tab), len(self.layers), self.any_norm, self.total_tiles, self.total_tiles2)
        if self.W3 is not None:
            call("artic_split", ptr(self.W3), ptr(self.W3_sp), self.x3_total, self.x3_total)

    def zero(self):
        self.dW.zero_()

    def unprep(self, grads: Dict[str, torch.Tensor]):
        """dW (prepared layout) -> dv / dg of the torch parameters (OVERWRITES those entries of
        ``grads``; bias gradients were accumulated by ConvLayer.wgrad)."""
        tab = self._table(grads)
        call("artic_weights_unprep", ptr(tab), len(self.layers), self.any_norm, self.total_tiles, self.total_tiles2)


def _zero_grads_like(layers: List[ConvLayer]) -> Dict[str, torch.Tensor]:
    out = {}
    for l in layers:
        for t, suffix in ((l.g, ".weight_g"), (l.v, ".weight_v" if l.g is not None else ".weight"), (l.b, ".bias")):
            if t is not None:
                out[l.name + suffix] = torch.zeros_like(t)
    return out


# --------------------------------------------------------------------------- #
# generator                                                                   #
# --------------------------------------------------------------------------- #

@contextlib.contextmanager
def planner_objective(sm_time_pct: int):
    """Tile-planner objective for the launches enqueued inside: percentage of the SM-time term against
    stand-alone latency (artic_debug_set key 15; 0 restores the default = 100).  The discriminator runs eight
    chains at once and wants small SM time per launch; the generator has three, so part of the machine is
    idle anyway and latency counts."""
    lib = _lib.load()
    lib.artic_debug_set(15, int(sm_time_pct))
    try:
        yield
    finally:
        lib.artic_debug_set(15, 0)


@contextlib.contextmanager
def weight_multicast(cluster_size: int):
    """Launches enqueued inside may run as thread-block clusters of ``cluster_size`` (2 or 4) CTAs that share every
    streamed weight tile through TMA multicast (artic_debug_set key 22).  Pays when few chains are in flight and the
    row count is small — the chunked-AR decoder: 19.6 -> 19.0 ms per 32 x 600-frame batch; inside the train step, where
    every SM is already held by some chain's CTA, co-scheduling CTA pairs costs more than the saved L2 traffic.
    An explicit ARTIC_DEBUG=22=... wins."""
    lib = _lib.load()
    if 22 in _lib.ENV_DEBUG:
        yield
        return
    _MCAST_STACK.append(int(cluster_size))
    lib.artic_debug_set(22, int(cluster_size))
    try:
        yield
    finally:
        _MCAST_STACK.pop()
        lib.artic_debug_set(22, _MCAST_STACK[-1] if _MCAST_STACK else 0)


_MCAST_STACK: List[int] = []
#: experiment knobs: cluster weight multicast for the generator's forward / backward launches inside the train step
_G_MCAST_FWD = int(_os.environ.get("ARTIC_G_MCAST_FWD", "0"))
_G_MCAST_BWD = int(_os.environ.get("ARTIC_G_MCAST_BWD", "0"))


#: ARTIC_FUSE_RES=0 turns the fused residual unit off (artic_resunit_fwd: conv1 -> LeakyReLU -> conv2 -> + x of the
#: narrow MRF stages in one launch; bf16 mode, C = 32 / 64)
_FUSE_RES = _os.environ.get("ARTIC_FUSE_RES", "1") != "0"
_FUSE_RES_BWD = _os.environ.get("ARTIC_FUSE_RES_BWD", "1") != "0"


def resunit_fusable(c1: "ConvLayer", c2: "ConvLayer") -> bool:
    """Both convs of a residual unit fit the fused kernel: bf16, C -> C with C in {32, 64}, same odd kernel size <= 11,
    conv2 undilated; the two weights stay resident in shared memory next to the tiles (C = 64, k = 11: 176 KB of weights,
    single-buffered tiles)."""
    s1, s2 = c1.spec, c2.spec
    C = s1.cin
    return (_FUSE_RES and c1.in_code == BF16 and c1.out_code == BF16 and c2.in_code == BF16 and c2.out_code == BF16
            and s1.kind == s2.kind == "conv" and s1.cout == C and s2.cin == C and s2.cout == C and C in (32, 64)
            and s1.groups == s2.groups == 1 and s1.stride == s2.stride == 1 and s1.k == s2.k and s1.k % 2 == 1 and s1.k <= 11
            and s2.dilation == 1 and s1.padding == (s1.k - 1) // 2 * s1.dilation and s2.padding == (s2.k - 1) // 2
            and (s1.k // 2) * s1.dilation <= 32 and c1.kcig == C and c2.kcig == C)


def resunit_forward(c1: "ConvLayer", c2: "ConvLayer", ax: SeqT, x: SeqT, at: Optional[SeqT], xn: Optional[SeqT],
                    axn: Optional[SeqT], slope: float):
    """One launch for at = lrelu(conv1(ax) + b1), xn = conv2(at) + b2 + x, axn = lrelu(xn) (any of at / xn / axn may be
    None, but one of xn / axn is required)."""
    assert ax.n_inner == 1 and ax.s_row == ax.C and ax.s_outer == ax.L * ax.C, "plain (N, L, C) batches only"
    p = _lib.ResUnit()
    p.AX, p.XRES, p.W1t, p.W2t = ptr(ax.t), ptr(x.t), ptr(c1.Wb), ptr(c2.Wb)
    p.b1, p.b2 = ptr(c1.b), ptr(c2.b)
    p.AT = ptr(at.t) if at is not None else None
    p.Y = ptr(xn.t) if xn is not None else None
    p.Y2 = ptr(axn.t) if axn is not None else None
    p.N, p.L, p.C, p.k, p.dil, p.slope = ax.N, ax.L, ax.C, c1.spec.k, c1.spec.dilation, slope
    call("artic_resunit_fwd", p)


def resunit_backward(c1: "ConvLayer", c2: "ConvLayer", gx: SeqT, at: SeqT, ax: SeqT, dt: SeqT, gn: SeqT, slope: float):
    """Data gradient of the unit in one launch: dt = conv2^T(gx) * lrelu'(at) (written out for conv1's weight
    gradient), gn = conv1^T(dt) * lrelu'(ax) + gx."""
    p = _lib.ResUnit()
    p.AX = p.XRES = ptr(gx.t)
    p.W1t, p.W2t = ptr(c2.Wf), ptr(c1.Wf)
    p.M1, p.M2 = ptr(at.t), ptr(ax.t)
    p.AT, p.Y = ptr(dt.t), ptr(gn.t)
    p.N, p.L, p.C, p.k, p.dil, p.slope, p.mode = gx.N, gx.L, gx.C, c1.spec.k, c1.spec.dilation, slope, 1
    call("artic_resunit_fwd", p)


_G_OBJECTIVE = int(_os.environ.get("ARTIC_G_OBJECTIVE", "60"))   # measured: 100 -> 12.93, 60 -> 12.82, 30 -> 13.15 ms
_D_OBJECTIVE = int(_os.environ.get("ARTIC_D_OBJECTIVE", "100"))  # the discriminator's eight chains: see profiles/r2_objective_sweep.log


def mlp_forward(x: torch.Tensor, lays, acts, code: int, slope: float):
    """Linear -> [LeakyReLU -> Linear] x (n - 1): x (B, dims[0]) fp32, acts[0] receives x in the storage type,
    acts[l + 1] layer l's output (activated for all but the last).  ONE launch (artic_mlp_fwd) when the widths allow,
    else one tap-conv launch per layer."""
    dims = [lays[0].spec.cin] + [lay.spec.cout for lay in lays]
    fused = len(lays) <= 8 and max(dims) <= 1024 and all(c % 8 == 0 and 1024 % (c // 8) == 0 for c in dims[1:]) and \
        all(lay.kcig == lay.spec.cin and lay.kcog == lay.spec.cout and lay.in_code == code for lay in lays)
    if not fused:
        call("artic_cast", ptr(x), F32, ptr(acts[0].t), code, x.numel())
        for l, lay in enumerate(lays):
            if l < len(lays) - 1:
                lay.forward(acts[l], Y2=acts[l + 1], act=ACT_LRELU, act_slope=slope)
            else:
                lay.forward(acts[l], Y=acts[l + 1])
        return
    p = _lib.Mlp()
    p.in_, p.act0 = ptr(x), ptr(acts[0].t)
    p.B, p.n_layers, p.dtype, p.slope = acts[0].N, len(lays), code, slope
    for l, d in enumerate(dims):
        p.dims[l] = d
    for l, lay in enumerate(lays):
        p.W[l], p.bias[l], p.outs[l] = ptr(lay.Wf), ptr(lay.b), ptr(acts[l + 1].t)
    call("artic_mlp_fwd", p)


class GeneratorEngine:
    """HiFiGANGenerator.forward (reference models/hifigan.py:198-239) and its backward."""

    def __init__(self, in_channels, out_channels, channels, kernel_size, upsample_scales,
                 upsample_kernel_sizes, paddings, output_paddings, resblock_kernel_sizes,
                 resblock_dilations, use_additional_convs, slope, use_weight_norm, use_ar, ar_input,
                 ar_hidden, ar_output, use_tanh, code=F32, x3=False):
        assert use_additional_convs, "use_additional_convs=False is not on the hot path"
        assert len(resblock_kernel_sizes) == 3, "the MRF mean kernel is written for 3 blocks"
        self.code, self.slope, self.use_ar, self.use_tanh = code, slope, use_ar, use_tanh
        self.x3 = bool(x3) and code == F32
        self.in_channels, self.out_channels = in_channels, out_channels
        self.ar_input, self.ar_output = ar_input, ar_output
        self.scales = list(upsample_scales)
        self.dilations = [list(d) for d in resblock_dilations]
        wn = use_weight_norm
        c = code
        L = self.layers = {}

        def add(name, spec, ic=c, oc=c, pad_in=False):
            spec.weight_norm = wn and spec.kind != "linear"
            spec.name = name
            L[name] = ConvLayer(spec, name, ic, oc, pad_in=pad_in, x3=self.x3)
            return L[name]

        add("input_conv", ConvSpec("conv", in_channels, channels, k=kernel_size, padding=(kernel_size - 1) // 2),
            pad_in=True)
        self.n_blocks = len(resblock_kernel_sizes)
        for i, (s, k) in enumerate(zip(upsample_scales, upsample_kernel_sizes)):
            add(f"upsamples.{i}.1", ConvSpec("convT", channels // 2 ** i, channels // 2 ** (i + 1), k=k, stride=s,
                                             padding=paddings[i], output_padding=output_paddings[i]))
            ch = channels // 2 ** (i + 1)
            for j, rk in enumerate(resblock_kernel_sizes):
                for di, d in enumerate(resblock_dilations[j]):
                    b = i * self.n_blocks + j
                    add(f"blocks.{b}.convs1.{di}.1", ConvSpec("conv", ch, ch, k=rk, dilation=d, padding=(rk - 1) // 2 * d))
                    add(f"blocks.{b}.convs2.{di}.1", ConvSpec("conv", ch, ch, k=rk, padding=(rk - 1) // 2))
        # the output conv reads a bf16/fp32 activation and writes the fp32 waveform
        add("output_conv.1", ConvSpec("conv", channels // 2 ** len(upsample_scales), out_channels, k=kernel_size,
                                      padding=(kernel_size - 1) // 2), ic=c, oc=F32)
        if use_ar:
            dims = [ar_input] + [ar_hidden] * 4 + [ar_output]
            for li in range(5):
                add(f"ar_model.model.{2 * li}", ConvSpec("linear", dims[li], dims[li + 1]))
        self._prepped = False

    # ---- parameters ----------------------------------------------------------
    def bind(self, params):
        for l in self.layers.values():
            l.bind(params)
        # keep the prepared-weight buffers and descriptor tables while the parameters stay in place
        # (re-binding happens after every optimizer update; CUDA-graph capture must not see uploads)
        sig = tuple(ptr(t) for l in self.layers.values() for t in (l.v, l.g, l.b))
        if getattr(self, "_bind_sig", None) != sig:
            self.wset = WeightSet(list(self.layers.values()))
            self._bind_sig = sig
        else:
            self.wset.rebind()
        self._prepped = False

    def prep_weights(self, need_bwd=True):
        self.wset.prep()
        self._prepped = True

    def param_names(self):
        return [n for l in self.layers.values() for n in l.param_names()]

    # ---- forward -------------------------------------------------------------
    def forward(self, c: torch.Tensor, ar: Optional[torch.Tensor], save=True):
        """c (B, Cc, T') fp32 channel-first, ar (B, 1, ar_input) fp32 -> ((B, 1, T) fp32, tape)."""
        with planner_objective(_G_OBJECTIVE):
            if _G_MCAST_FWD:
                with weight_multicast(_G_MCAST_FWD):
                    return self._forward(c, ar, save)
            return self._forward(c, ar, save)

    def _forward(self, c, ar, save):
        _lib.require_cuda(c, "c")
        assert self._prepped, "call prep_weights() after binding / updating parameters"
        L, code, dev, slope = self.layers, self.code, c.device, self.slope
        B, Cc, Tn = c.shape
        c = c.contiguous().float()
        tape = {"B": B, "Tn": Tn, "Cc": Cc} if save else None
        ar_feats = None
        Ca = 0
        if self.use_ar:
            Ca = self.ar_output
            a0 = SeqT.empty(B, 1, self.ar_input, code, dev)
            lays = [L[f"ar_model.model.{2 * li}"] for li in range(5)]
            acts = [a0] + [SeqT.empty(B, 1, lay.spec.cout, code, dev) for lay in lays]
            mlp_forward(ar.contiguous().float(), lays, acts, code, 0.1)          # pytorch_layers.py:440,446
            h = acts[-1]
            ar_feats = h
            if save:
                tape["ar_acts"] = acts
        assert Cc + Ca == self.in_channels, f"in_channels {self.in_channels} != {Cc} + {Ca}"
        Cpad = L["input_conv"].kcig                      # >= Cc + Ca (zero-padded for the tensor-core path)
        gin = SeqT.empty(B, Tn, Cpad, code, dev)
        call("artic_gen_input", ptr(c), ptr(ar_feats.t) if ar_feats is not None else None, ptr(gin.t),
             B, Cc, Ca, Cpad, Tn, code)
        a = SeqT.empty(B, Tn, L["input_conv"].spec.cout, code, dev)
        L["input_conv"].forward(gin, Y2=a, act=ACT_LRELU, act_slope=slope)
        if save:
            tape["gin"], tape["stages"] = gin, []
        n_stage = len(self.scales)
        for i in range(n_stage):
            up = L[f"upsamples.{i}.1"]
            lo = up.spec.out_len(a.L)
            u = SeqT.empty(B, lo, up.spec.cout, code, dev)
            au = u.like()
            up.forward(a, Y=u, Y2=au, act=ACT_LRELU, act_slope=slope)
            st = {"a_in": a, "a_u": au, "blocks": []}

            def block_fwd(j, i=i, u=u, au=au):
                b = i * self.n_blocks + j
                x, ax = u, au
                pairs = []
                nd = len(self.dilations[j])
                for di in range(nd):
                    c1, c2 = L[f"blocks.{b}.convs1.{di}.1"], L[f"blocks.{b}.convs2.{di}.1"]
                    xn = u.like()
                    axn = u.like() if di < nd - 1 else None
                    if resunit_fusable(c1, c2):
                        # narrow stages: both convs in one launch, the intermediate stays on chip (it is written out
                        # only when the backward will need it)
                        at = u.like() if save else None
                        resunit_forward(c1, c2, ax, x, at, xn, axn, slope)
                    else:
                        at = u.like()
                        c1.forward(ax, Y2=at, act=ACT_LRELU, act_slope=slope)
                        c2.forward(at, Y=xn, Y2=axn, res=x, act=ACT_LRELU, act_slope=slope)
                    pairs.append((ax, at))
                    x, ax = xn, axn
                return x, pairs

            if _GROUPED:
                # the three MRF blocks advance in lock-step: one multi-problem launch per conv depth
                nb = self.n_blocks
                nd = len(self.dilations[0])
                assert all(len(d) == nd for d in self.dilations)
                xs, axs = [u] * nb, [au] * nb
                pairs_all = [[] for _ in range(nb)]
                for di in range(nd):
                    ats = [u.like() for _ in range(nb)]
                    prm = []
                    for j in range(nb):
                        prm += L[f"blocks.{i * nb + j}.convs1.{di}.1"].forward_params(axs[j], Y2=ats[j], act=ACT_LRELU, act_slope=slope)
                    launch_tapconvs(prm)
                    xns = [u.like() for _ in range(nb)]
                    axns = [u.like() if di < nd - 1 else None for _ in range(nb)]
                    prm = []
                    for j in range(nb):
                        prm += L[f"blocks.{i * nb + j}.convs2.{di}.1"].forward_params(ats[j], Y=xns[j], Y2=axns[j], res=xs[j],
                                                                                 act=ACT_LRELU, act_slope=slope)
                    launch_tapconvs(prm)
                    for j in range(nb):
                        pairs_all[j].append((axs[j], ats[j]))
                    xs, axs = xns, axns
                outs = xs
                st["blocks"] = pairs_all
            else:
                res_j = fork_join([lambda j=j: block_fwd(j) for j in range(self.n_blocks)])
                outs = [r[0] for r in res_j]
                st["blocks"] = [r[1] for r in res_j]
            last = i == n_stage - 1
            # LeakyReLU before the output conv uses torch's default slope 0.01 (hifigan.py:150)
            st["slope_out"] = 0.01 if last else slope
            a = u.like()
            call("artic_mean3_act", ptr(outs[0].t), ptr(outs[1].t), ptr(outs[2].t), ptr(a.t), a.numel(),
                 st["slope_out"], code, code)
            st["a_c"] = a
            if save:
                tape["stages"].append(st)
        y = SeqT.empty(B, a.L, self.out_channels, F32, dev)
        if self.use_tanh:
            L["output_conv.1"].forward(a, Y2=y, act=ACT_TANH)
        else:
            L["output_conv.1"].forward(a, Y=y)
        if save:
            tape["y"] = y
        # (B, T, C_out) channels-last -> (B, C_out, T); zero-copy for out_channels == 1
        out = y.t.permute(0, 2, 1) if self.out_channels > 1 else y.t.view(B, 1, a.L)
        return out, tape

    # ---- backward ------------------------------------------------------------
    def backward(self, tape, dy: torch.Tensor, grads: Dict[str, torch.Tensor]):
        """dy: (B, C_out, T) fp32 gradient of the waveform.  Accumulates parameter gradients
        into ``grads`` (name -> fp32 tensor shaped like the parameter)."""
        with planner_objective(_G_OBJECTIVE):
            if _G_MCAST_BWD:
                with weight_multicast(_G_MCAST_BWD):
                    return self._backward(tape, dy, grads)
            return self._backward(tape, dy, grads)

    def _backward(self, tape, dy, grads):
        L, code, slope = self.layers, self.code, self.slope
        B = tape["B"]
        dev = dy.device
        self.wset.zero()
        y = tape["y"]
        dyc = dy.permute(0, 2, 1).contiguous().float() if self.out_channels > 1 else dy.contiguous().float()
        dpre = SeqT.empty(B, y.L, self.out_channels, F32, dev)
        if self.use_tanh:
            call("artic_tanh_bwd", ptr(dyc), ptr(y.t), ptr(dpre.t), dpre.numel(), F32)
        else:
            dpre.t.view(-1).copy_(dyc.reshape(-1))
        stages = tape["stages"]
        oc = L["output_conv.1"]
        a_c = stages[-1]["a_c"]
        oc.wgrad(a_c, dpre, grads)
        # gradient wrt each MRF block output of the last stage: (1/3) * lrelu'(a_c) * dgrad
        g = a_c.like()
        oc.dgrad(dpre, dX=g, mask=a_c, mask_slope=stages[-1]["slope_out"], alpha=1.0 / self.n_blocks, sp_y=self.x3)
        for i in range(len(stages) - 1, -1, -1):
            st = stages[i]

            def block_bwd(j, i=i, st=st, g=g):
                b = i * self.n_blocks + j
                gx = g
                pairs = st["blocks"][j]
                wq = SideQueue()                 # weight gradients run beside the data-gradient chain
                for di in range(len(pairs) - 1, -1, -1):
                    ax, at = pairs[di]
                    c2, c1 = L[f"blocks.{b}.convs2.{di}.1"], L[f"blocks.{b}.convs1.{di}.1"]
                    wq.run(lambda c2=c2, at=at, gx=gx: c2.wgrad(at, gx, grads), at, gx)
                    dt = at.like()
                    gn = ax.like()
                    if _FUSE_RES_BWD and resunit_fusable(c1, c2):
                        resunit_backward(c1, c2, gx, at, ax, dt, gn, slope)      # both data gradients, one launch
                        wq.run(lambda c1=c1, ax=ax, dt=dt: c1.wgrad(ax, dt, grads), ax, dt)
                    else:
                        c2.dgrad(gx, dX=dt, mask=at, mask_slope=slope)
                        wq.run(lambda c1=c1, ax=ax, dt=dt: c1.wgrad(ax, dt, grads), ax, dt)
                        c1.dgrad(dt, dX=gn, mask=ax, mask_slope=slope, res=gx)
                    gx = gn
                return gx, wq        # gradient wrt the block input (pre-activation u); queue joined by the caller

            if _GROUPED:
                nb = self.n_blocks
                gxs = [g] * nb
                wq = SideQueue()                 # weight gradients run beside the data-gradient chain
                for di in range(len(st["blocks"][0]) - 1, -1, -1):
                    prs = [st["blocks"][j][di] for j in range(nb)]          # (ax, at) of every block
                    c2s = [L[f"blocks.{i * nb + j}.convs2.{di}.1"] for j in range(nb)]
                    c1s = [L[f"blocks.{i * nb + j}.convs1.{di}.1"] for j in range(nb)]
                    for j in range(nb):
                        wq.run(lambda c2=c2s[j], at=prs[j][1], gx=gxs[j]: c2.wgrad(at, gx, grads), prs[j][1], gxs[j])
                    dts = [prs[j][1].like() for j in range(nb)]
                    prm = []
                    for j in range(nb):
                        prm += c2s[j].dgrad_params(gxs[j], dX=dts[j], mask=prs[j][1], mask_slope=slope)
                    launch_tapconvs(prm)
                    for j in range(nb):
                        wq.run(lambda c1=c1s[j], ax=prs[j][0], dt=dts[j]: c1.wgrad(ax, dt, grads), prs[j][0], dts[j])
                    gns = [prs[j][0].like() for j in range(nb)]
                    prm = []
                    for j in range(nb):
                        prm += c1s[j].dgrad_params(dts[j], dX=gns[j], mask=prs[j][0], mask_slope=slope, res=gxs[j])
                    launch_tapconvs(prm)
                    gxs = gns
                dus = gxs
                held = wq.join()                 # noqa: F841
            else:
                res_b = fork_join([lambda j=j: block_bwd(j) for j in range(self.n_blocks)])
                dus = [r[0] for r in res_b]
                held = [r[1].join() for r in res_b]      # noqa: F841  (tensors of queued wgrads stay alive until joined)
            du = dus[0].like()
            call("artic_sum3", ptr(dus[0].t), ptr(dus[1].t), ptr(dus[2].t), ptr(du.t), du.numel(), code)
            if self.x3:
                du.ensure_split()                # before the weight gradient's side stream forks off
            up = L[f"upsamples.{i}.1"]
            a_in = st["a_in"]
            up.wgrad(a_in, du, grads)
            g = a_in.like()
            if i > 0:
                up.dgrad(du, dX=g, mask=a_in, mask_slope=stages[i - 1]["slope_out"], alpha=1.0 / self.n_blocks)
            else:
                up.dgrad(du, dX=g, mask=a_in, mask_slope=slope)
        ic = L["input_conv"]
        gin = tape["gin"]
        ic.wgrad(gin, g, grads)
        if self.use_ar:
            dgin = gin.like()
            ic.dgrad(g, dX=dgin, sp_y=False)
            Ca = self.ar_output
            d_ar32 = torch.empty((B, Ca), dtype=torch.float32, device=dev)
            call("artic_gen_input_bwd", ptr(dgin.t), ptr(d_ar32), B, tape["Cc"], Ca, gin.C, tape["Tn"], code)
            dz = SeqT.empty(B, 1, Ca, code, dev)
            call("artic_cast", ptr(d_ar32), F32, ptr(dz.t), code, B * Ca)
            acts = tape["ar_acts"]
            for li in range(4, -1, -1):
                lay = L[f"ar_model.model.{2 * li}"]
                lay.wgrad(acts[li], dz, grads)
                if li > 0:
                    dn = acts[li].like()
                    lay.dgrad(dz, dX=dn, mask=acts[li], mask_slope=0.1)
                    dz = dn
        self.wset.unprep(grads)

    def new_grads(self):
        return _zero_grads_like(list(self.layers.values()))


# --------------------------------------------------------------------------- #
# discriminator                                                               #
# --------------------------------------------------------------------------- #
class _Chain:
    """One sub-discriminator: a chain of conv layers, LeakyReLU after all but the last."""

    def __init__(self, layers: List[ConvLayer], kind: str, period: int = 1, scale_index: int = 0):
        self.layers, self.kind, self.period, self.scale_index = layers, kind, period, scale_index


class DiscriminatorEngine:
    """HiFiGANMultiScaleMultiPeriodDiscriminator (reference models/hifigan.py:741-825):
    ``scales`` scale discriminators on an AvgPool1d pyramid followed by one period
    discriminator per period.  Feature maps are stored in ``code`` dtype, the first conv of
    every chain reads the fp32 signal and the logits are written in fp32."""

    def __init__(self, scales, pool_params, scale_params, follow_official_norm, periods, period_params, code=F32,
                 scale_prefix="msd.discriminators.{i}.", period_prefix="mpd.discriminators.{i}.", x3=False):
        """``scale_prefix`` / ``period_prefix`` name the parameters of sub-discriminator ``i`` (the
        stand-alone classes use "discriminators.{i}." or ""); ``scales`` may be 0 and ``periods`` empty."""
        from .convspec import ConvSpec as CS
        self.code = code
        self.x3 = bool(x3) and code == F32
        self.pool = dict(pool_params or {"kernel_size": 4, "stride": 2, "padding": 2})
        scale_params = scale_params or {}
        period_params = period_params or {}
        self.slope_s = scale_params.get("nonlinear_activation_params", {}).get("negative_slope", 0.1)
        self.slope_p = period_params.get("nonlinear_activation_params", {}).get("negative_slope", 0.1)
        self.chains: List[_Chain] = []
        sp, pp = scale_params, period_params
        # NOTE reference quirk: the MSD norm hooks test isinstance(m, Conv2d) on
        # Conv1d layers (hifigan.py:645-663), so NO weight/spectral norm is ever applied to MSD.
        for s in range(scales):
            ks = sp["kernel_sizes"]
            specs = [CS("conv", sp["in_channels"], sp["channels"], k=ks[0], padding=(ks[0] - 1) // 2)]
            in_chs = out_chs = sp["channels"]
            groups = 4
            for ds in sp["downsample_scales"]:
                specs.append(CS("conv", in_chs, out_chs, k=ks[1], stride=ds, padding=(ks[1] - 1) // 2, groups=groups))
                in_chs = out_chs
                out_chs = min(in_chs * 2, sp["max_downsample_channels"])
                groups = min(groups * 4, sp["max_groups"])
            out_chs = min(in_chs * 2, sp["max_downsample_channels"])
            specs.append(CS("conv", in_chs, out_chs, k=ks[2], padding=(ks[2] - 1) // 2))
            specs.append(CS("conv", out_chs, sp["out_channels"], k=ks[3], padding=(ks[3] - 1) // 2))
            layers = []
            for li, spec in enumerate(specs):
                last = li == len(specs) - 1
                name = scale_prefix.format(i=s) + f"layers.{li}" + ("" if last else ".0")
                layers.append(ConvLayer(spec, name, F32 if li == 0 else code, F32 if last else code, x3=self.x3))
            self.chains.append(_Chain(layers, "scale", scale_index=s))
        for pi, period in enumerate(periods):
            ks = pp["kernel_sizes"]
            specs = []
            in_chs, out_chs = pp["in_channels"], pp["channels"]
            for ds in pp["downsample_scales"]:
                specs.append(CS("conv", in_chs, out_chs, k=ks[0], stride=ds, padding=(ks[0] - 1) // 2))
                in_chs = out_chs
                out_chs = min(out_chs * 4, pp["max_downsample_channels"])
            # reference: Conv2d(out_chs, out, (ks[1]-1, 1), 1, padding=((ks[1]-1)//2, 0)), hifigan.py:382-388
            specs.append(CS("conv", out_chs, pp["out_channels"], k=ks[1] - 1, padding=(ks[1] - 1) // 2))
            assert out_chs == in_chs, "output conv width must equal the last feature width (shipped configs)"
            layers = []
            for li, spec in enumerate(specs):
                last = li == len(specs) - 1
                name = period_prefix.format(i=pi) + ("output_conv" if last else f"convs.{li}.0")
                layers.append(ConvLayer(spec, name, F32 if li == 0 else code, F32 if last else code, x3=self.x3))
            self.chains.append(_Chain(layers, "period", period=period))
        self.layers = {l.name: l for ch in self.chains for l in ch.layers}
        self._flat_specs = {}
        self._prepped = False

    def bind(self, params):
        for l in self.layers.values():
            l.bind(params)
        # keep the prepared-weight buffers and descriptor tables while the parameters stay in place
        # (re-binding happens after every optimizer update; CUDA-graph capture must not see uploads)
        sig = tuple(ptr(t) for l in self.layers.values() for t in (l.v, l.g, l.b))
        if getattr(self, "_bind_sig", None) != sig:
            self.wset = WeightSet(list(self.layers.values()))
            self._bind_sig = sig
        else:
            self.wset.rebind()
        self._prepped = False

    def prep_weights(self, need_bwd=True):
        self.wset.prep()
        self._prepped = True

    def param_names(self):
        return [n for l in self.layers.values() for n in l.param_names()]

    def new_grads(self):
        return _zero_grads_like(list(self.layers.values()))

    def _flat_spec(self, ch, lay):
        """Dilated-conv restatement of a stride-1 period layer (see flat_period), or None."""
        s = lay.spec
        if ch.kind != "period" or s.stride != 1 or s.k == 1 or s.kind != "conv" or s.cin < 32 or s.cout < 32 \
                or _os.environ.get("ARTIC_FLAT_PERIOD", "1") == "0":
            return None
        key = (lay.name, ch.period)
        fs = self._flat_specs.get(key)
        if fs is None:
            from .convspec import ConvSpec as CS
            fs = CS("conv", s.cin, s.cout, k=s.k, dilation=ch.period, padding=s.padding * ch.period, groups=s.groups)
            self._flat_specs[key] = fs
        return fs

    # ---- forward -------------------------------------------------------------
    def forward(self, x: Optional[torch.Tensor], save=True, into=None, lo=0, parts=None):
        with planner_objective(_D_OBJECTIVE):
            return self._forward(x, save, into, lo, parts)

    def _forward(self, x: Optional[torch.Tensor], save=True, into=None, lo=0, parts=None):
        """x (B, 1, T) fp32 -> (list of 8 lists of SeqT [feature maps..., logits], tape).

        ``parts = (ar, ys)`` instead of ``x``: the input is cat([ar, y], dim=2) for every y in ``ys`` (one or two
        (B0, 1, Ty) batches) stacked along the batch (bin/train.py:345-346) — assembled by the same single launch
        that builds the pooled and reflect-padded views.

        ``into`` (a tape of an earlier forward over a batch of >= lo + B items) makes this call
        write its activations into batch items [lo, lo + B) of that tape's buffers instead of
        allocating: the train step keeps [fake | real] in ONE 2B batch so that the discriminator
        backward (and every weight gradient) runs once over both halves."""
        assert self._prepped
        if parts is not None:
            ar, ys = parts
            B0, _, Ty = ys[0].shape
            La = 0 if ar is None else ar.shape[-1]
            B, T = len(ys) * B0, La + Ty
            dev = ys[0].device
            for t in ys:
                _lib.require_cuda(t, "y")
                assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (B0, 1, Ty)
            assert ar is None or (ar.dtype == torch.float32 and ar.is_contiguous() and ar.numel() == B0 * La)
        else:
            _lib.require_cuda(x, "x")
            assert x.dim() == 3 and x.shape[1] == 1
            B, _, T = x.shape
            dev = x.device
        if into is not None:
            assert into["T"] == T and lo + B <= into["B"]
            x2d = into["x"][lo:lo + B]
        elif parts is not None:
            x2d = torch.empty((B, T), dtype=torch.float32, device=dev)
        else:
            x2d = x.contiguous().float().view(B, T)

        def take(full: SeqT):
            return slice_seq(full, lo, lo + B)

        outs, tape = [], {"B": B, "T": T, "x": x2d, "chains": [], "xp": {}}
        # signal assembly + AvgPool pyramid (hifigan.py:733-736) + reflect-padded period inputs (hifigan.py:413-416):
        # one launch (artic_disc_prep)
        k, st, pd = self.pool["kernel_size"], self.pool["stride"], self.pool["padding"]
        sigs = [SeqT(x2d.view(B, T, 1), B, T, 1)]
        n_scales = sum(1 for c in self.chains if c.kind == "scale")
        for s in range(1, n_scales):
            lo_ = (sigs[-1].L + 2 * pd - k) // st + 1
            sigs.append(take(into["sigs"][s]) if into is not None else SeqT.empty(B, lo_, 1, F32, dev))
        xps, padded = {}, []
        for ci, ch in enumerate(self.chains):
            if ch.kind != "period":
                continue
            p = ch.period
            Tp = T if T % p == 0 else T + (p - T % p)
            if Tp != T:
                xp = into["xp"][ci][lo:lo + B] if into is not None else torch.empty((B, Tp), dtype=torch.float32, device=dev)
                padded.append(xp)
            else:
                xp = x2d
            xps[ci] = xp
        tape["xp"] = xps
        dp = _lib.DiscPrep()
        if parts is not None:
            dp.ar = ptr(ar)
            for i, t in enumerate(ys):
                dp.y[i] = ptr(t)
            dp.B, dp.La, dp.x_out = B0, La, ptr(x2d)
        else:
            xin = x2d if into is None else x.contiguous().float().view(B, T)
            dp.x, dp.x_out = ptr(xin), (ptr(x2d) if into is not None else None)
        dp.N, dp.T, dp.k, dp.stride, dp.pad = B, T, k, st, pd
        assert len(sigs) - 1 <= 4 and len(padded) <= 8
        dp.n_pool, dp.n_xp = len(sigs) - 1, len(padded)
        for i, sg in enumerate(sigs[1:]):
            dp.pool[i], dp.pool_len[i] = ptr(sg.t), sg.L
        for i, xp in enumerate(padded):
            dp.xp[i], dp.xp_len[i] = ptr(xp), xp.shape[1]
        if 4 * (T + sum(sg.L for sg in sigs[1:])) > _DISC_PREP_MAX_BYTES:
            # a row with its pyramid does not fit the one-launch kernel's shared memory: the single-purpose kernels
            if parts is not None:
                for i, t in enumerate(ys):
                    call("artic_concat_time", ptr(ar) if La else None, ptr(t), ptr(x2d[i * B0:]), B0, La, Ty, T, F32)
            elif into is not None:
                x2d.copy_(x.reshape(B, T))
            for a, b in zip(sigs[:-1], sigs[1:]):
                call("artic_avgpool1d", ptr(a.t), ptr(b.t), B, a.L, b.L, k, st, pd, F32)
            for xp in padded:
                call("artic_reflect_pad_right", ptr(x2d), ptr(xp), B, T, xp.shape[1], F32)
        elif dp.n_pool or dp.n_xp or dp.x_out:
            call("artic_disc_prep", dp)

        def chain_fwd(ci):
            ch = self.chains[ci]
            if ch.kind == "scale":
                h = sigs[ch.scale_index]
                slope = self.slope_s
            else:
                p = ch.period
                xp = xps[ci]
                Tp = xp.shape[1]
                h = SeqT(xp, B * p, Tp // p, 1, n_inner=p, s_outer=Tp, s_inner=1, s_row=p)
                slope = self.slope_p
            acts = [h]
            n = len(ch.layers)
            for li, lay in enumerate(ch.layers):
                lo_ = lay.spec.out_len(h.L)
                last = li == n - 1
                if into is not None:
                    o = take(into["chains"][ci][li + 1])
                else:
                    o = h.like(code=F32 if last else self.code, C=lay.spec.cout, L=lo_)
                fs = self._flat_spec(ch, lay)
                if last:
                    lay.forward(h, Y=o)
                elif fs is not None:
                    launch_tapconvs(lay.forward_params(flat_period(h), Y2=flat_period(o), act=ACT_LRELU, act_slope=slope, spec=fs,
                                                       sp_y2=self.x3))
                else:
                    lay.forward(h, Y2=o, act=ACT_LRELU, act_slope=slope, sp_y2=self.x3)
                acts.append(o)
                h = o
            return acts

        all_acts = fork_join([lambda ci=ci: chain_fwd(ci) for ci in range(len(self.chains))])
        for acts in all_acts:
            outs.append(acts[1:])
            tape["chains"].append(acts)
        tape["sigs"] = sigs
        return outs, (tape if save else None)

    @staticmethod
    def slice_tape(tape, lo, hi):
        """The tape of batch items [lo, hi) of a larger forward (views, no copies)."""
        return {"B": hi - lo, "T": tape["T"], "x": tape["x"][lo:hi],
                "xp": {ci: t[lo:hi] for ci, t in tape["xp"].items()},
                "sigs": [slice_seq(s, lo, hi) for s in tape["sigs"]],
                "chains": [[slice_seq(a, lo, hi) for a in acts] for acts in tape["chains"]]}

    # ---- backward ------------------------------------------------------------
    def backward(self, tape, douts, grads: Optional[Dict[str, torch.Tensor]], need_dx=True, pre_zeroed=False, finalize=True):
        with planner_objective(_D_OBJECTIVE):
            return self._backward(tape, douts, grads, need_dx, pre_zeroed, finalize)

    def _backward(self, tape, douts, grads: Optional[Dict[str, torch.Tensor]], need_dx=True, pre_zeroed=False, finalize=True):
        """douts: per chain a list (same length as the chain's outputs) of SeqT gradients or
        None; the gradient wrt the logits must be present.  Accumulates parameter gradients
        into ``grads`` when given (None = skip every wgrad, as in the generator phase) and
        returns d x (B, 1, T) fp32 when ``need_dx``.  ``pre_zeroed``: the caller already cleared the
        weight-gradient accumulators (``wset.zero()``) off the critical path.  ``finalize=False``: leave the weight
        gradients in the prepared-layout accumulators (another backward over other batch items will add to them and
        finish with ``wset.unprep``)."""
        B, T = tape["B"], tape["T"]
        dev = tape["sigs"][0].t.device
        if grads is not None and not pre_zeroed:
            self.wset.zero()
        dx = torch.zeros((B, 1, T), dtype=torch.float32, device=dev) if need_dx else None
        n_scales = len(tape["sigs"])
        dsig = [None] * n_scales

        def chain_bwd(ci):
            ch = self.chains[ci]
            acts = tape["chains"][ci]           # acts[0] = input signal, acts[l+1] = output of layer l
            dl = douts[ci]
            slope = self.slope_s if ch.kind == "scale" else self.slope_p
            n = len(ch.layers)
            dz = dl[n - 1]                      # logits gradient (fp32 SeqT)
            assert dz is not None
            d_in = None
            wq = SideQueue() if grads is not None else None
            for li in range(n - 1, -1, -1):
                lay = ch.layers[li]
                fs = self._flat_spec(ch, lay) if li < n - 1 else None
                if grads is not None:
                    if fs is not None:
                        wq.run(lambda lay=lay, a=acts[li], dz=dz, fs=fs: lay.wgrad(flat_period(a), flat_period(dz), grads, spec=fs),
                               acts[li], dz)
                    else:
                        wq.run(lambda lay=lay, a=acts[li], dz=dz: lay.wgrad(a, dz, grads), acts[li], dz)
                if li > 0:
                    dn = acts[li].like()
                    if fs is not None:
                        rp = flat_period(dl[li - 1]) if dl[li - 1] is not None else None
                        launch_tapconvs(lay.dgrad_params(flat_period(dz), dX=flat_period(dn), res_pre=rp, mask=flat_period(acts[li]),
                                                         mask_slope=slope, spec=fs, sp_y=self.x3))
                    else:
                        lay.dgrad(dz, dX=dn, res_pre=dl[li - 1], mask=acts[li], mask_slope=slope, sp_y=self.x3)
                    dz = dn
                elif need_dx:
                    if ch.kind == "scale":
                        d_in = acts[0].like(code=F32)
                        lay.dgrad(dz, dX=d_in)
                    else:
                        h = acts[0]
                        Tp = h.s_outer
                        dxp = torch.empty((B, Tp), dtype=torch.float32, device=dev)
                        d_in = SeqT(dxp, h.N, h.L, 1, n_inner=h.n_inner, s_outer=Tp, s_inner=1, s_row=h.s_row)
                        lay.dgrad(dz, dX=d_in)
            return d_in, wq

        res_c = fork_join([lambda ci=ci: chain_bwd(ci) for ci in range(len(self.chains))])
        d_ins = [r[0] for r in res_c]
        held = [r[1].join() for r in res_c if r[1] is not None]      # noqa: F841
        if need_dx:     # the chains' input gradients meet in dx on the main stream
            for ch, d_in in zip(self.chains, d_ins):
                if ch.kind == "scale":
                    dsig[ch.scale_index] = d_in
                else:
                    call("artic_reflect_pad_right_bwd", ptr(d_in.t), ptr(dx), B, T, d_in.s_outer, 1, F32)
        if need_dx:
            k, st, pd = self.pool["kernel_size"], self.pool["stride"], self.pool["padding"]
            for s in range(n_scales - 1, 0, -1):
                lp = tape["sigs"][s - 1].L
                call("artic_avgpool1d_bwd", ptr(dsig[s].t), ptr(dsig[s - 1].t), B, lp, tape["sigs"][s].L,
                     k, st, pd, 1, F32)
            # dx += dsig[0]  (via the reflect-pad backward with Lp == L: plain accumulate)
            call("artic_reflect_pad_right_bwd", ptr(dsig[0].t), ptr(dx), B, T, T, 1, F32)
        if grads is not None and finalize:
            self.wset.unprep(grads)
        return dx
