"""HiFi-GAN / HiFi-CAR generator and discriminators — the plugin surface.

Drop-in for reference articulatory/models/hifigan.py: same class names (resolved from
YAML by ``getattr(models, config["generator_type"])``, reference bin/train.py:1649-1662),
same constructor keywords, same ``state_dict`` keys (weight-norm ``weight_g`` /
``weight_v`` pairs on the generator and the period discriminators, plain weights on the
scale discriminators), same ``forward`` / ``inference`` / ``remove_weight_norm`` /
``apply_weight_norm`` / ``register_stats`` methods.

The torch submodules below are PARAMETER CONTAINERS only (their own ``forward`` is never
called): all compute goes through ``engine.GeneratorEngine`` / ``DiscriminatorEngine``,
i.e. the sm_100a kernels of libartic_sm100.so.  Inputs must live on a CUDA device; there
is no CPU fallback.
"""
import copy
import logging

import numpy as np
import torch

from .. import _lib
from ..engine import DiscriminatorEngine, GeneratorEngine, SeqT

#: storage dtype of hidden activations per precision mode.  "fp32": fp32 storage, CUDA-core kernels (debug / odd
#: shapes).  "bf16x3": fp32 storage, every eligible contraction on the tcgen05 kernels as an error-compensated
#: split product (x_hi w_hi + x_hi w_lo + x_lo w_hi, fp32 accumulate) — the PARITY-GATED tensor-core mode (1e-3 on
#: waveforms / losses against the fp32 reference).  "bf16": bf16 storage, plain tcgen05 (speed mode).
_PRECISIONS = {"fp32": _lib.F32, "bf16x3": _lib.F32, "bf16": _lib.BF16}
#: default storage precision of hidden activations for newly built models
DEFAULT_PRECISION = "fp32"


def set_default_precision(name):
    global DEFAULT_PRECISION
    assert name in _PRECISIONS
    DEFAULT_PRECISION = name


def _wn(module):
    return torch.nn.utils.weight_norm(module)


def _params_version(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class _EngineModule(torch.nn.Module):
    """Common weight-materialisation cache: prepared weights are rebuilt only when a
    parameter changed (torch bumps ``_version`` on every in-place update; the fused Adam
    of this package calls ``mark_weights_dirty``)."""

    def _init_engine_state(self):
        self._engine = None
        self._prep_key = None
        self._dirty = True

    def mark_weights_dirty(self):
        self._dirty = True

    def _named_param_tensors(self):
        return {n: p.data for n, p in self.named_parameters()}

    def _ensure_ready(self):
        key = _params_version(self)
        if self._engine is None or self._dirty or self._prep_key != key:
            if self._engine is None:
                self._engine = self._build_engine()
            self._engine.bind(self._named_param_tensors())
            self._engine.prep_weights(need_bwd=True)
            self._prep_key = key
            self._dirty = False
        return self._engine

    def _apply(self, fn, *a, **k):  # .to() / .cuda() re-allocate parameters
        self._engine = None
        self._prep_key = None
        return super()._apply(fn, *a, **k)

    def set_precision(self, name):
        assert name in _PRECISIONS
        self.precision = name
        self._engine = None
        self._prep_key = None


# --------------------------------------------------------------------------- #
# generator                                                                   #
# --------------------------------------------------------------------------- #
class _PastFC(torch.nn.Module):
    """Parameter container of the CAR past-sample encoder (reference
    layers/pytorch_layers.py:426-449): ``model.{0,2,4,6,8}`` are the Linear layers."""

    def __init__(self, input_len, hidden_dim, output_dim):
        super().__init__()
        mods = [torch.nn.Linear(input_len, hidden_dim), torch.nn.LeakyReLU(0.1)]
        for _ in range(3):
            mods += [torch.nn.Linear(hidden_dim, hidden_dim), torch.nn.LeakyReLU(0.1)]
        mods.append(torch.nn.Linear(hidden_dim, output_dim))
        self.model = torch.nn.Sequential(*mods)


class _ResBlockParams(torch.nn.Module):
    """Parameter container of one MRF block (reference layers/residual_block.py:172-205):
    ``convs1.{d}.1`` dilated conv, ``convs2.{d}.1`` plain conv."""

    def __init__(self, kernel_size, channels, dilations, bias, use_additional_convs, slope):
        super().__init__()
        self.convs1 = torch.nn.ModuleList()
        if use_additional_convs:
            self.convs2 = torch.nn.ModuleList()
        for d in dilations:
            self.convs1.append(torch.nn.Sequential(
                torch.nn.LeakyReLU(slope),
                torch.nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, bias=bias,
                                padding=(kernel_size - 1) // 2 * d)))
            if use_additional_convs:
                self.convs2.append(torch.nn.Sequential(
                    torch.nn.LeakyReLU(slope),
                    torch.nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, bias=bias,
                                    padding=(kernel_size - 1) // 2)))


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, need_grad, c, ar, *params):
        eng = mod._ensure_ready()
        out, tape = eng.forward(c, ar, save=need_grad)
        ctx.mod, ctx.tape, ctx.eng = mod, tape, eng
        ctx.names = [n for n, _ in mod.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        grads = eng.new_grads()
        eng.backward(ctx.tape, dy, grads)
        ctx.tape = None
        return (None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


class HiFiGANGenerator(_EngineModule):
    """HiFiGAN generator with optional CAR conditioning (reference models/hifigan.py:21-314)."""

    def __init__(self, in_channels=80, out_channels=1, channels=512, kernel_size=7,
                 upsample_scales=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4), paddings=None,
                 output_paddings=None, resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilations=[(1, 3, 5), (1, 3, 5), (1, 3, 5)], use_additional_convs=True, bias=True,
                 nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
                 use_weight_norm=True, use_ar=False, ar_input=512, ar_hidden=256, ar_output=128, use_tanh=True,
                 use_spk_id=False, num_spk=None, spk_emb_size=32, use_ph=False, num_ph=None, ph_emb_size=8,
                 use_ph_loss=False,
                 # keys present in egs/ema/voc1/conf/e2w_hifigan_car.yaml:42,54 that the reference
                 # constructor rejects: accepted and ignored so the yaml runs unchanged
                 final_scale=None, extra_art=None,
                 precision=None):
        super().__init__()
        if use_spk_id or use_ph or use_ph_loss:
            raise NotImplementedError("speaker / phoneme conditioning is outside the B200 hot path (SURVEY.md §2.4)")
        if nonlinear_activation != "LeakyReLU":
            raise NotImplementedError("only LeakyReLU is implemented on the B200 path")
        assert kernel_size % 2 == 1, "Kernel size must be odd number."
        assert len(upsample_scales) == len(upsample_kernel_sizes)
        assert len(resblock_dilations) == len(resblock_kernel_sizes)
        self.use_ar = use_ar
        self.precision = DEFAULT_PRECISION if precision is None else precision
        slope = nonlinear_activation_params.get("negative_slope", 0.01)

        def rule(vals, default):
            if vals is None:
                return [default(s) for s in upsample_scales]
            out = []
            for s, v in zip(upsample_scales, vals):
                if v != "default":
                    raise NotImplementedError("only 'default' paddings are supported (as in the reference)")
                out.append(default(s))
            return out

        paddings = rule(paddings, lambda s: s // 2 + s % 2)              # reference hifigan.py:82-92
        output_paddings = rule(output_paddings, lambda s: s % 2)          # :93-103
        self._cfg = dict(in_channels=in_channels, out_channels=out_channels, channels=channels,
                         kernel_size=kernel_size, upsample_scales=list(upsample_scales),
                         upsample_kernel_sizes=list(upsample_kernel_sizes), paddings=paddings,
                         output_paddings=output_paddings, resblock_kernel_sizes=list(resblock_kernel_sizes),
                         resblock_dilations=[list(d) for d in resblock_dilations],
                         use_additional_convs=use_additional_convs, slope=slope, use_weight_norm=use_weight_norm,
                         use_ar=use_ar, ar_input=ar_input, ar_hidden=ar_hidden, ar_output=ar_output,
                         use_tanh=use_tanh)

        # ---- parameter containers, created in the reference's order so that the same
        # torch seed yields the same initial weights ----
        self.num_upsamples = len(upsample_kernel_sizes)
        self.num_blocks = len(resblock_kernel_sizes)
        self.input_conv = torch.nn.Conv1d(in_channels, channels, kernel_size, 1, padding=(kernel_size - 1) // 2)
        self.upsamples = torch.nn.ModuleList()
        self.blocks = torch.nn.ModuleList()
        for i in range(self.num_upsamples):
            self.upsamples.append(torch.nn.Sequential(
                torch.nn.LeakyReLU(slope),
                torch.nn.ConvTranspose1d(channels // 2 ** i, channels // 2 ** (i + 1), upsample_kernel_sizes[i],
                                         upsample_scales[i], padding=paddings[i], output_padding=output_paddings[i])))
            for j in range(self.num_blocks):
                self.blocks.append(_ResBlockParams(resblock_kernel_sizes[j], channels // 2 ** (i + 1),
                                                   resblock_dilations[j], bias, use_additional_convs, slope))
        tail = [torch.nn.LeakyReLU(),   # default slope 0.01, as in the reference (:150)
                torch.nn.Conv1d(channels // 2 ** self.num_upsamples, out_channels, kernel_size, 1,
                                padding=(kernel_size - 1) // 2)]
        if use_tanh:
            tail.append(torch.nn.Tanh())
        self.output_conv = torch.nn.Sequential(*tail)
        if use_ar:
            self.ar_model = _PastFC(ar_input, ar_hidden, ar_output)
        if use_weight_norm:
            self.apply_weight_norm()
        self.reset_parameters()
        self._init_engine_state()

    # ---- reference-compatible utilities --------------------------------------------
    def reset_parameters(self):
        """The reference (hifigan.py:241-254) draws N(0, 0.01) into ``m.weight`` AFTER weight
        norm was applied, which leaves weight_g / weight_v untouched but advances the RNG
        (SURVEY.md §7 quirks).  Reproduced: same draws, same (non-)effect."""
        for m in self.modules():
            if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
                m.weight.data.normal_(0.0, 0.01)

    def apply_weight_norm(self):
        for m in list(self.modules()):
            if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)) and not hasattr(m, "weight_g"):
                _wn(m)
        self._invalidate()

    def remove_weight_norm(self):
        for m in list(self.modules()):
            try:
                torch.nn.utils.remove_weight_norm(m)
            except ValueError:
                pass
        self._invalidate()

    def _invalidate(self):
        if hasattr(self, "_engine"):
            self._engine = None
            self._prep_key = None

    def register_stats(self, stats):
        """Register (mean, scale) for input normalisation (reference hifigan.py:280-296)."""
        assert stats.endswith(".h5") or stats.endswith(".npy")
        if stats.endswith(".h5"):
            import h5py  # optional dependency, same as the reference
            with h5py.File(stats, "r") as f:
                mean, scale = f["mean"][()].reshape(-1), f["scale"][()].reshape(-1)
        else:
            mean, scale = np.load(stats)[0].reshape(-1), np.load(stats)[1].reshape(-1)
        self.register_buffer("mean", torch.from_numpy(mean).float())
        self.register_buffer("scale", torch.from_numpy(scale).float())
        logging.info("Successfully registered stats as buffer.")

    # ---- compute -------------------------------------------------------------------
    def _build_engine(self):
        wn = any(n.endswith("weight_g") for n, _ in self.named_parameters())
        cfg = dict(self._cfg, use_weight_norm=wn)
        return GeneratorEngine(code=_PRECISIONS[self.precision], x3=self.precision == "bf16x3", **cfg)

    def forward(self, c, spk_id=None, ar=None, ph=None):
        """c (B, in_channels - ar_output, T') [, ar (B, 1, ar_input)] -> (B, out_channels, T'*prod(scales))."""
        if self.use_ar and ar is None:
            raise ValueError("use_ar=True requires the `ar` argument")
        params = [p for _, p in self.named_parameters()]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _GenFn.apply(self, need_grad, c, ar if self.use_ar else None, *params)

    def inference(self, c, normalize_before=False):
        """(T', in_channels) -> (T, out_channels) (reference hifigan.py:298-314; like the
        reference it does not pass ``ar`` and therefore only serves use_ar=False models —
        CAR models are decoded with ``articulatory_b200.bin.decode.ar_loop``)."""
        if not isinstance(c, torch.Tensor):
            c = torch.tensor(c, dtype=torch.float).to(next(self.parameters()).device)
        if normalize_before:
            c = (c - self.mean) / self.scale
        with torch.no_grad():
            c = self.forward(c.transpose(1, 0).unsqueeze(0))
        return c.squeeze(0).transpose(1, 0)


# --------------------------------------------------------------------------- #
# discriminators                                                              #
# --------------------------------------------------------------------------- #
def _seq_to_ref_view(o: SeqT, kind: str, is_logits: bool):
    """Engine storage -> tensor shaped like the reference output (zero-copy views)."""
    if kind == "scale":
        return o.t.permute(0, 2, 1)                                   # (B, L, C) -> (B, C, L)
    B = o.N // o.n_inner
    t = o.t                                                           # (B, H, p, C)
    if is_logits:
        return t.permute(0, 3, 1, 2).flatten(1, -1)                   # == torch.flatten(x, 1, -1), hifigan.py:425
    return t.permute(0, 3, 1, 2)                                      # (B, C, H, p)


def _grad_to_seq(g, ref: SeqT, kind: str, is_logits: bool):
    """Inverse of _seq_to_ref_view for an incoming gradient (None stays None)."""
    if g is None:
        return None
    if kind == "scale":
        t = g.permute(0, 2, 1)
    else:
        B = ref.N // ref.n_inner
        if is_logits:
            g = g.reshape(B, ref.C, ref.L, ref.n_inner)
        t = g.permute(0, 2, 3, 1)
    t = t.to(ref.t.dtype).contiguous()
    return SeqT(t, ref.N, ref.L, ref.C, ref.n_inner, ref.s_outer, ref.s_inner, ref.s_row)


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, need_w, need_x, x, *params):
        eng = mod._ensure_ready()
        outs, tape = eng.forward(x, save=need_w or need_x)
        flat, meta = [], []
        for ci, (ch, lst) in enumerate(zip(eng.chains, outs)):
            for li, o in enumerate(lst):
                flat.append(_seq_to_ref_view(o, ch.kind, li == len(lst) - 1))
                meta.append((ci, li, ch.kind, li == len(lst) - 1, o))
        ctx.mod, ctx.eng, ctx.tape, ctx.meta = mod, eng, tape, meta
        ctx.need_w, ctx.need_x = need_w, need_x
        ctx.names = [n for n, _ in mod.named_parameters()]
        return tuple(flat)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.eng
        douts = [[None] * len(ch.layers) for ch in eng.chains]
        for g, (ci, li, kind, is_logits, ref) in zip(gouts, ctx.meta):
            if g is None and is_logits:
                g = torch.zeros_like(_seq_to_ref_view(ref, kind, True))
            douts[ci][li] = _grad_to_seq(g, ref, kind, is_logits)
        grads = eng.new_grads() if ctx.need_w else None
        dx = eng.backward(ctx.tape, douts, grads, need_dx=ctx.need_x)
        ctx.tape = None
        pg = tuple(grads.get(n) for n in ctx.names) if grads is not None else tuple(None for _ in ctx.names)
        return (None, None, None, dx) + pg


_DEFAULT_SCALE_PARAMS = {
    "in_channels": 1, "out_channels": 1, "kernel_sizes": [15, 41, 5, 3], "channels": 128,
    "max_downsample_channels": 1024, "max_groups": 16, "bias": True,
    "downsample_scales": [2, 2, 4, 4, 1], "nonlinear_activation": "LeakyReLU",
    "nonlinear_activation_params": {"negative_slope": 0.1}}
_DEFAULT_PERIOD_PARAMS = {
    "in_channels": 1, "out_channels": 1, "kernel_sizes": [5, 3], "channels": 32,
    "downsample_scales": [3, 3, 3, 3, 1], "max_downsample_channels": 1024, "bias": True,
    "nonlinear_activation": "LeakyReLU", "nonlinear_activation_params": {"negative_slope": 0.1},
    "use_weight_norm": True, "use_spectral_norm": False}
_DEFAULT_POOL_PARAMS = {"kernel_size": 4, "stride": 2, "padding": 2}


class _DiscriminatorBase(_EngineModule):
    """Shared body of the five discriminator plugin classes: validates the reference keywords,
    builds the engine topology once on the host to enumerate the layers, and mirrors it with
    torch parameter containers named like the reference's modules (``layers.N[.0]``,
    ``convs.N.0``, ``output_conv``, under ``scale_prefix`` / ``period_prefix``)."""

    def _setup(self, scales, pool, pool_params, scale_params, follow_official_norm, periods, period_params,
               scale_prefix, period_prefix, precision):
        if pool != "AvgPool1d":
            raise NotImplementedError("only AvgPool1d pooling is implemented")
        scale_params = copy.deepcopy(scale_params) if scale_params else None
        period_params = copy.deepcopy(period_params) if period_params else None
        if period_params is not None:
            period_params = dict(_DEFAULT_PERIOD_PARAMS, **period_params)
            period_params.pop("period", None)
            if period_params.get("use_spectral_norm", False):
                raise NotImplementedError("spectral norm is not on the hot path (yaml: use_spectral_norm false)")
            ks = period_params["kernel_sizes"]
            assert len(ks) == 2 and ks[0] % 2 == 1 and ks[1] % 2 == 1, "Kernel size must be odd number."
        if scale_params is not None:
            scale_params = dict(_DEFAULT_SCALE_PARAMS, **scale_params)
            ks = scale_params["kernel_sizes"]
            assert len(ks) == 4 and all(k % 2 == 1 for k in ks)
        for prm in (scale_params, period_params):
            if prm is None:
                continue
            if prm.get("nonlinear_activation", "LeakyReLU") != "LeakyReLU":
                raise NotImplementedError("only LeakyReLU is implemented on the B200 path")
            assert prm.get("bias", True), "bias=False is not on the hot path"
        self.precision = DEFAULT_PRECISION if precision is None else precision
        self._cfg = dict(scales=scales, pool_params=copy.deepcopy(pool_params), scale_params=scale_params,
                         follow_official_norm=follow_official_norm, periods=list(periods),
                         period_params=period_params, scale_prefix=scale_prefix, period_prefix=period_prefix)
        topo = DiscriminatorEngine(code=_lib.F32, **self._cfg)
        use_wn_p = (period_params or {}).get("use_weight_norm", True)
        for ch in topo.chains:
            n = len(ch.layers)
            for li, lay in enumerate(ch.layers):
                s = lay.spec
                if ch.kind == "scale":
                    # reference quirk (hifigan.py:645-663): the norm hooks test isinstance(m, Conv2d) on
                    # Conv1d layers, so no weight / spectral norm is ever applied to a scale discriminator
                    conv = torch.nn.Conv1d(s.cin, s.cout, s.k, stride=s.stride, padding=s.padding, groups=s.groups)
                else:
                    conv = torch.nn.Conv2d(s.cin, s.cout, (s.k, 1), (s.stride, 1), padding=(s.padding, 0))
                    if use_wn_p:
                        conv = _wn(conv)
                self._place(lay.name, conv)
        self._init_engine_state()

    def _place(self, dotted, module):
        """Register ``module`` under a dotted path, creating plain container modules on the way."""
        parts = dotted.split(".")
        node = self
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, torch.nn.Module())
            node = node._modules[part]
        node.add_module(parts[-1], module)

    def _build_engine(self):
        return DiscriminatorEngine(code=_PRECISIONS[self.precision], x3=self.precision == "bf16x3", **self._cfg)

    def _forward_lists(self, x):
        params = [p for _, p in self.named_parameters()]
        need_w = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        need_x = torch.is_grad_enabled() and x.requires_grad
        flat = _DiscFn.apply(self, need_w, need_x, x, *params)
        outs, i = [], 0
        for ch in self._engine.chains:
            n = len(ch.layers)
            outs.append(list(flat[i:i + n]))
            i += n
        return outs

    def forward(self, x):
        return self._forward_lists(x)


class HiFiGANPeriodDiscriminator(_DiscriminatorBase):
    """HiFiGAN period discriminator (reference models/hifigan.py:317-440): ``forward(x (B,1,T))`` returns
    the list of the 5 feature maps (B, C, T/p', p) followed by the flattened logits."""

    def __init__(self, in_channels=1, out_channels=1, period=3, kernel_sizes=[5, 3], channels=32,
                 downsample_scales=[3, 3, 3, 3, 1], max_downsample_channels=1024, bias=True,
                 nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
                 use_weight_norm=True, use_spectral_norm=False, precision=None):
        super().__init__()
        if use_weight_norm and use_spectral_norm:
            raise ValueError("Either use use_weight_norm or use_spectral_norm.")
        self.period = period
        pp = dict(in_channels=in_channels, out_channels=out_channels, kernel_sizes=kernel_sizes, channels=channels,
                  downsample_scales=downsample_scales, max_downsample_channels=max_downsample_channels, bias=bias,
                  nonlinear_activation=nonlinear_activation, nonlinear_activation_params=nonlinear_activation_params,
                  use_weight_norm=use_weight_norm, use_spectral_norm=use_spectral_norm)
        self._setup(0, "AvgPool1d", None, None, False, [period], pp, "", "", precision)

    def forward(self, x):
        return self._forward_lists(x)[0]


class HiFiGANMultiPeriodDiscriminator(_DiscriminatorBase):
    """HiFiGAN multi-period discriminator (reference models/hifigan.py:443-500): list of per-period lists."""

    def __init__(self, periods=[2, 3, 5, 7, 11], discriminator_params=_DEFAULT_PERIOD_PARAMS, precision=None):
        super().__init__()
        self._setup(0, "AvgPool1d", None, None, False, periods, discriminator_params, "", "discriminators.{i}.",
                    precision)


class HiFiGANScaleDiscriminator(_DiscriminatorBase):
    """HiFi-GAN scale discriminator (reference models/hifigan.py:503-663): list of the 7 feature maps
    followed by the logits (B, 1, T')."""

    def __init__(self, in_channels=1, out_channels=1, kernel_sizes=[15, 41, 5, 3], channels=128,
                 max_downsample_channels=1024, max_groups=16, bias=True, downsample_scales=[2, 2, 4, 4, 1],
                 nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
                 use_weight_norm=True, use_spectral_norm=False, precision=None):
        super().__init__()
        if use_weight_norm and use_spectral_norm:
            raise ValueError("Either use use_weight_norm or use_spectral_norm.")
        sp = dict(in_channels=in_channels, out_channels=out_channels, kernel_sizes=kernel_sizes, channels=channels,
                  max_downsample_channels=max_downsample_channels, max_groups=max_groups, bias=bias,
                  downsample_scales=downsample_scales, nonlinear_activation=nonlinear_activation,
                  nonlinear_activation_params=nonlinear_activation_params)
        self._setup(1, "AvgPool1d", None, sp, False, [], None, "", "", precision)

    def forward(self, x):
        return self._forward_lists(x)[0]


class HiFiGANMultiScaleDiscriminator(_DiscriminatorBase):
    """HiFi-GAN multi-scale discriminator (reference models/hifigan.py:666-738): list of per-scale lists,
    scale i seeing the input pooled i times."""

    def __init__(self, scales=3, downsample_pooling="AvgPool1d", downsample_pooling_params=_DEFAULT_POOL_PARAMS,
                 discriminator_params=_DEFAULT_SCALE_PARAMS, follow_official_norm=False, precision=None):
        super().__init__()
        self._setup(scales, downsample_pooling, downsample_pooling_params, discriminator_params,
                    follow_official_norm, [], None, "discriminators.{i}.", "", precision)


class HiFiGANMultiScaleMultiPeriodDiscriminator(_DiscriminatorBase):
    """HiFi-GAN multi-scale + multi-period discriminator (reference models/hifigan.py:741-825).

    ``forward(x (B,1,T))`` returns 8 lists (3 scales, then 5 periods); each inner list holds
    the per-layer feature maps followed by the logits, with the reference's shapes."""

    def __init__(self, scales=3, scale_downsample_pooling="AvgPool1d",
                 scale_downsample_pooling_params=_DEFAULT_POOL_PARAMS,
                 scale_discriminator_params=_DEFAULT_SCALE_PARAMS,
                 follow_official_norm=True, periods=[2, 3, 5, 7, 11],
                 period_discriminator_params=_DEFAULT_PERIOD_PARAMS,
                 precision=None):
        super().__init__()
        self._setup(scales, scale_downsample_pooling, scale_downsample_pooling_params, scale_discriminator_params,
                    follow_official_norm, periods, period_discriminator_params,
                    "msd.discriminators.{i}.", "mpd.discriminators.{i}.", precision)
