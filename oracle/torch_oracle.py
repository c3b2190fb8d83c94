"""CPU oracle: torch-fp32/fp64 restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY — never imported by ``articulatory_b200``.

Every function cites the reference file:line it restates (paths relative to
/root/reference/articulatory).  The oracle is functional: it operates on a plain
``state_dict`` whose keys are the reference's own parameter names, so the same
weights can be fed to the reference (when importable), the oracle and the CUDA
product.  It is pinned against the unmodified reference in
``tests/test_oracle_vs_reference.py`` (runs only where /root/reference exists) and
against the committed fixtures in ``tests/golden/`` (generated from the reference by
``tests/golden/make_golden.py``).

Floating point: all arithmetic is done in the dtype of the inputs (fp32 to
mirror the reference; pass float64 tensors for a high-precision run).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle.mel_basis import slaney_mel_basis

# --------------------------------------------------------------------------- #
# configuration helpers                                                       #
# --------------------------------------------------------------------------- #

#: egs/ema/voc1/conf/e2w_hifigan.yaml:33-56 (generator_params)
E2W_GENERATOR_PARAMS = dict(
    in_channels=141, out_channels=1, channels=512, kernel_size=7,
    upsample_scales=[5, 4, 2, 2], upsample_kernel_sizes=[10, 8, 4, 4],
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilations=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    use_additional_convs=True, bias=True, nonlinear_activation="LeakyReLU",
    nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True,
    use_ar=True, ar_input=512, ar_hidden=256, ar_output=128,
)

#: egs/ema/voc1/conf/e2w_hifigan.yaml:61-95 (discriminator_params)
E2W_DISCRIMINATOR_PARAMS = dict(
    scales=3, scale_downsample_pooling="AvgPool1d",
    scale_downsample_pooling_params=dict(kernel_size=4, stride=2, padding=2),
    scale_discriminator_params=dict(
        in_channels=1, out_channels=1, kernel_sizes=[15, 41, 5, 3], channels=128,
        max_downsample_channels=1024, max_groups=16, bias=True,
        downsample_scales=[4, 4, 4, 4, 1], nonlinear_activation="LeakyReLU",
        nonlinear_activation_params={"negative_slope": 0.1}),
    follow_official_norm=True, periods=[2, 3, 5, 7, 11],
    period_discriminator_params=dict(
        in_channels=1, out_channels=1, kernel_sizes=[5, 3], channels=32,
        downsample_scales=[3, 3, 3, 3, 1], max_downsample_channels=1024, bias=True,
        nonlinear_activation="LeakyReLU",
        nonlinear_activation_params={"negative_slope": 0.1},
        use_weight_norm=True, use_spectral_norm=False),
)

#: egs/ema/voc1/conf/e2w_hifigan.yaml:102-111
E2W_MEL_LOSS_PARAMS = dict(fs=16000, fft_size=1024, hop_size=80, win_length=None,
                           window="hann", num_mels=80, fmin=0, fmax=11025, log_base=None)

#: losses/stft_loss.py:124-129 (defaults of MultiResolutionSTFTLoss)
DEFAULT_STFT_LOSS_PARAMS = dict(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50],
                                win_lengths=[600, 1200, 240], window="hann_window")


def _gen_defaults(p):
    """Fill the defaults of HiFiGANGenerator.__init__ (models/hifigan.py:24-50)."""
    d = dict(in_channels=80, out_channels=1, channels=512, kernel_size=7,
             upsample_scales=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4),
             paddings=None, output_paddings=None, resblock_kernel_sizes=(3, 7, 11),
             resblock_dilations=[(1, 3, 5)] * 3, use_additional_convs=True, bias=True,
             nonlinear_activation="LeakyReLU",
             nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True,
             use_ar=False, ar_input=512, ar_hidden=256, ar_output=128, use_tanh=True)
    d.update({k: v for k, v in p.items() if k not in ("final_scale", "extra_art")})
    s = d["upsample_scales"]
    # models/hifigan.py:82-103: default padding rule
    if d["paddings"] is None:
        d["paddings"] = [x // 2 + x % 2 for x in s]
    if d["output_paddings"] is None:
        d["output_paddings"] = [x % 2 for x in s]
    return d


# --------------------------------------------------------------------------- #
# weights                                                                     #
# --------------------------------------------------------------------------- #

def weight_norm_weight(g, v):
    """w = g * v / ||v||, norm over all dims but 0 (torch.nn.utils.weight_norm,
    default dim=0; applied at models/hifigan.py:268-278 and :430-438)."""
    dims = tuple(range(1, v.dim()))
    return g * v / v.pow(2).sum(dim=dims, keepdim=True).sqrt()


def effective_weight(sd, prefix):
    """Conv weight for module ``prefix``: plain ``weight`` or weight-normed pair."""
    if prefix + ".weight_g" in sd:
        return weight_norm_weight(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])
    return sd[prefix + ".weight"]


def init_generator_state(params, seed=0, dtype=torch.float32):
    """Random generator state_dict with the reference's key names and shapes
    (models/hifigan.py:105-196; layers/residual_block.py:172-205;
    layers/pytorch_layers.py:437-449).  Values: N(0, 0.02)-ish — used only where the
    reference itself cannot be constructed (GPU box)."""
    p = _gen_defaults(params)
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, co, ci, k, transposed=False):
        shape = (ci, co, k) if transposed else (co, ci, k)
        v = torch.randn(shape, generator=g, dtype=dtype) * (1.0 / math.sqrt(ci * k))
        if p["use_weight_norm"]:
            sd[prefix + ".bias"] = torch.randn(co, generator=g, dtype=dtype) * 0.01
            sd[prefix + ".weight_g"] = v.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
            sd[prefix + ".weight_v"] = v
        else:
            sd[prefix + ".weight"] = v
            sd[prefix + ".bias"] = torch.randn(co, generator=g, dtype=dtype) * 0.01

    ch = p["channels"]
    conv("input_conv", ch, p["in_channels"], p["kernel_size"])
    nb = len(p["resblock_kernel_sizes"])
    for i, (s, k) in enumerate(zip(p["upsample_scales"], p["upsample_kernel_sizes"])):
        conv(f"upsamples.{i}.1", ch // 2 ** (i + 1), ch // 2 ** i, k, transposed=True)
        for j, rk in enumerate(p["resblock_kernel_sizes"]):
            c = ch // 2 ** (i + 1)
            for d in range(len(p["resblock_dilations"][j])):
                conv(f"blocks.{i * nb + j}.convs1.{d}.1", c, c, rk)
                if p["use_additional_convs"]:
                    conv(f"blocks.{i * nb + j}.convs2.{d}.1", c, c, rk)
    conv("output_conv.1", p["out_channels"], ch // 2 ** len(p["upsample_scales"]), p["kernel_size"])
    if p["use_ar"]:
        dims = [p["ar_input"]] + [p["ar_hidden"]] * 4 + [p["ar_output"]]
        for li in range(5):
            fan_in = dims[li]
            sd[f"ar_model.model.{2 * li}.weight"] = (
                torch.rand(dims[li + 1], fan_in, generator=g, dtype=dtype) * 2 - 1) / math.sqrt(fan_in)
            sd[f"ar_model.model.{2 * li}.bias"] = (
                torch.rand(dims[li + 1], generator=g, dtype=dtype) * 2 - 1) / math.sqrt(fan_in)
    return sd


def init_discriminator_state(params, seed=1, dtype=torch.float32):
    """Random MSMPD state_dict with the reference's key names and shapes
    (models/hifigan.py:361-388, :549-615).  MSD convs are plain ``weight``/``bias``
    (the reference's MSD norm hooks never match Conv1d, models/hifigan.py:645-663);
    MPD convs are weight-normed Conv2d (k,1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sp = params["scale_discriminator_params"]
    pp = params["period_discriminator_params"]

    def rnd(shape, fan_in):
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) / math.sqrt(fan_in)

    for s in range(params["scales"]):
        for li, (co, cig, k, _st, _g) in enumerate(msd_layer_specs(sp)):
            pre = f"msd.discriminators.{s}.layers.{li}" + (".0" if li < len(msd_layer_specs(sp)) - 1 else "")
            sd[pre + ".weight"] = rnd((co, cig, k), cig * k)
            sd[pre + ".bias"] = rnd((co,), cig * k)
    for pi, _period in enumerate(params["periods"]):
        specs = mpd_layer_specs(pp)
        for li, (co, ci, k, _st) in enumerate(specs):
            last = li == len(specs) - 1
            pre = f"mpd.discriminators.{pi}." + ("output_conv" if last else f"convs.{li}.0")
            v = rnd((co, ci, k, 1), ci * k)
            sd[pre + ".bias"] = rnd((co,), ci * k)
            if pp.get("use_weight_norm", True):
                sd[pre + ".weight_g"] = v.pow(2).sum(dim=(1, 2, 3), keepdim=True).sqrt()
                sd[pre + ".weight_v"] = v
            else:
                sd[pre + ".weight"] = v
    return sd


def msd_layer_specs(sp):
    """[(C_out, C_in/groups, k, stride, groups)] per layer, models/hifigan.py:549-615."""
    ks = sp["kernel_sizes"]
    specs = [(sp["channels"], sp["in_channels"], ks[0], 1, 1)]
    in_chs = out_chs = sp["channels"]
    groups = 4
    for ds in sp["downsample_scales"]:
        specs.append((out_chs, in_chs // groups, ks[1], ds, groups))
        in_chs = out_chs
        out_chs = min(in_chs * 2, sp["max_downsample_channels"])
        groups = min(groups * 4, sp["max_groups"])
    out_chs = min(in_chs * 2, sp["max_downsample_channels"])
    specs.append((out_chs, in_chs, ks[2], 1, 1))
    specs.append((sp["out_channels"], out_chs, ks[3], 1, 1))
    return specs


def mpd_layer_specs(pp):
    """[(C_out, C_in, k, stride)] per layer incl. output conv, models/hifigan.py:361-388."""
    ks = pp["kernel_sizes"]
    specs = []
    in_chs, out_chs = pp["in_channels"], pp["channels"]
    for ds in pp["downsample_scales"]:
        specs.append((out_chs, in_chs, ks[0], ds))
        in_chs = out_chs
        out_chs = min(out_chs * 4, pp["max_downsample_channels"])
    # NOTE reference quirk: output conv takes ``out_chs`` (already multiplied) as its
    # input width; with the shipped config both are 1024 (models/hifigan.py:381-388).
    specs.append((pp["out_channels"], out_chs, ks[1] - 1, 1))
    return specs


# --------------------------------------------------------------------------- #
# generator                                                                   #
# --------------------------------------------------------------------------- #

def past_fc_encoder(sd, ar, prefix="ar_model.model"):
    """PastFCEncoder.forward (layers/pytorch_layers.py:451-460): 5 Linear layers,
    LeakyReLU(0.1) after the first four."""
    x = ar.reshape(ar.shape[0], -1)
    for li in range(5):
        x = F.linear(x, sd[f"{prefix}.{2 * li}.weight"], sd[f"{prefix}.{2 * li}.bias"])
        if li < 4:
            x = F.leaky_relu(x, 0.1)
    return x


def residual_block(sd, prefix, x, kernel_size, dilations, slope, use_additional_convs=True):
    """HiFiGANResidualBlock.forward (layers/residual_block.py:207-222)."""
    for idx, d in enumerate(dilations):
        xt = F.conv1d(F.leaky_relu(x, slope), effective_weight(sd, f"{prefix}.convs1.{idx}.1"),
                      sd.get(f"{prefix}.convs1.{idx}.1.bias"), padding=(kernel_size - 1) // 2 * d, dilation=d)
        if use_additional_convs:
            xt = F.conv1d(F.leaky_relu(xt, slope), effective_weight(sd, f"{prefix}.convs2.{idx}.1"),
                          sd.get(f"{prefix}.convs2.{idx}.1.bias"), padding=(kernel_size - 1) // 2)
        x = xt + x
    return x


def generator_forward(sd, params, c, ar=None, return_intermediates=False):
    """HiFiGANGenerator.forward (models/hifigan.py:198-239), spk/ph branches excluded."""
    p = _gen_defaults(params)
    slope = p["nonlinear_activation_params"]["negative_slope"]
    inter = {}
    if p["use_ar"]:
        ar_feats = past_fc_encoder(sd, ar)                                # :209
        ar_feats = ar_feats.unsqueeze(2).repeat(1, 1, c.shape[2])         # :210
        c = torch.cat((c, ar_feats), dim=1)                               # :211
    k = p["kernel_size"]
    c = F.conv1d(c, effective_weight(sd, "input_conv"), sd["input_conv.bias"], padding=(k - 1) // 2)
    inter["input_conv"] = c
    nb = len(p["resblock_kernel_sizes"])
    for i, s in enumerate(p["upsample_scales"]):
        c = F.conv_transpose1d(F.leaky_relu(c, slope), effective_weight(sd, f"upsamples.{i}.1"),
                               sd[f"upsamples.{i}.1.bias"], stride=s, padding=p["paddings"][i],
                               output_padding=p["output_paddings"][i])    # :119-133
        inter[f"upsample{i}"] = c
        cs = 0.0
        for j in range(nb):
            cs = cs + residual_block(sd, f"blocks.{i * nb + j}", c, p["resblock_kernel_sizes"][j],
                                     p["resblock_dilations"][j], slope, p["use_additional_convs"])
        c = cs / nb                                                       # :226-230
        inter[f"mrf{i}"] = c
    # output conv: LeakyReLU with torch's DEFAULT slope 0.01 (:150), conv, tanh
    c = F.conv1d(F.leaky_relu(c, 0.01), effective_weight(sd, "output_conv.1"), sd["output_conv.1.bias"],
                 padding=(k - 1) // 2)
    if p["use_tanh"]:
        c = torch.tanh(c)
    return (c, inter) if return_intermediates else c


# --------------------------------------------------------------------------- #
# discriminators                                                              #
# --------------------------------------------------------------------------- #

def mpd_padded_length(t, period):
    """Reflect-pad rule of HiFiGANPeriodDiscriminator.forward (models/hifigan.py:412-417)."""
    return t if t % period == 0 else t + (period - t % period)


def period_discriminator_forward(sd, prefix, pp, period, x):
    """HiFiGANPeriodDiscriminator.forward (models/hifigan.py:401-428)."""
    slope = pp["nonlinear_activation_params"]["negative_slope"]
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t += n_pad
    x = x.view(b, c, t // period, period)
    specs = mpd_layer_specs(pp)
    outs = []
    for li, (co, ci, k, st) in enumerate(specs[:-1]):
        pre = f"{prefix}.convs.{li}.0"
        x = F.conv2d(x, effective_weight(sd, pre), sd[pre + ".bias"], stride=(st, 1), padding=((k - 1) // 2, 0))
        x = F.leaky_relu(x, slope)
        outs.append(x)
    pre = f"{prefix}.output_conv"
    ks1 = pp["kernel_sizes"][1]
    x = F.conv2d(x, effective_weight(sd, pre), sd[pre + ".bias"], stride=1, padding=((ks1 - 1) // 2, 0))
    outs.append(torch.flatten(x, 1, -1))
    return outs


def scale_discriminator_forward(sd, prefix, sp, x):
    """HiFiGANScaleDiscriminator.forward (models/hifigan.py:628-643)."""
    slope = sp["nonlinear_activation_params"]["negative_slope"]
    specs = msd_layer_specs(sp)
    outs = []
    for li, (co, cig, k, st, g) in enumerate(specs):
        last = li == len(specs) - 1
        pre = f"{prefix}.layers.{li}" + ("" if last else ".0")
        x = F.conv1d(x, effective_weight(sd, pre), sd.get(pre + ".bias"), stride=st, padding=(k - 1) // 2, groups=g)
        if not last:
            x = F.leaky_relu(x, slope)
        outs.append(x)
    return outs


def discriminator_forward(sd, params, x):
    """HiFiGANMultiScaleMultiPeriodDiscriminator.forward (models/hifigan.py:811-825):
    MSD outputs (with AvgPool1d between scales, :733-736) followed by MPD outputs."""
    outs = []
    pool = params["scale_downsample_pooling_params"]
    xs = x
    for s in range(params["scales"]):
        outs.append(scale_discriminator_forward(sd, f"msd.discriminators.{s}", params["scale_discriminator_params"], xs))
        xs = F.avg_pool1d(xs, pool["kernel_size"], pool["stride"], pool["padding"])
    for pi, period in enumerate(params["periods"]):
        outs.append(period_discriminator_forward(sd, f"mpd.discriminators.{pi}", params["period_discriminator_params"], period, x))
    return outs


# --------------------------------------------------------------------------- #
# losses                                                                      #
# --------------------------------------------------------------------------- #

def stft_frames(t, hop):
    """#frames of torch.stft(center=True): 1 + T // hop (losses/stft_loss.py:30-33)."""
    return 1 + t // hop


def stft_magnitude(x, fft_size, hop_size, win_length, window):
    """stft() of losses/stft_loss.py:16-40 → (B, frames, fft_size//2+1)."""
    # (.float(): a no-op in fp32; under bf16 autocast — bench.py's stock-torch GPU leg — cuFFT needs fp32 input, which
    # is what torch.autocast's own fp32 list does for the other spectral ops)
    s = torch.stft(x.float() if x.dtype == torch.bfloat16 else x, fft_size, hop_size, win_length, window, return_complex=True)
    power = s.real ** 2 + s.imag ** 2
    return torch.sqrt(torch.clamp(power, min=1e-7)).transpose(2, 1)


def mr_stft_loss(x, y, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50),
                 win_lengths=(600, 1200, 240), window="hann_window"):
    """MultiResolutionSTFTLoss.forward (losses/stft_loss.py:146-170): returns
    (spectral convergence, log STFT magnitude), each the mean over resolutions."""
    if x.dim() == 3:
        x = x.reshape(-1, x.size(2))
        y = y.reshape(-1, y.size(2))
    sc, mag = 0.0, 0.0
    for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths):
        w = getattr(torch, window)(wl, dtype=x.dtype, device=x.device)
        xm = stft_magnitude(x, fs, ss, wl, w)
        ym = stft_magnitude(y, fs, ss, wl, w)
        sc = sc + torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")      # :61
        mag = mag + F.l1_loss(torch.log(ym), torch.log(xm))                    # :82
    n = len(fft_sizes)
    return sc / n, mag / n


def mel_spectrogram(x, fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann",
                    num_mels=80, fmin=80, fmax=7600, center=True, normalized=False, onesided=True,
                    eps=1e-10, log_base=10.0):
    """MelSpectrogram.forward (losses/mel_loss.py:82-111) → (B, mels, frames)."""
    if x.dim() == 3:
        x = x.reshape(-1, x.size(2))
    win_length = fft_size if win_length is None else win_length
    w = getattr(torch, f"{window}_window")(win_length, dtype=x.dtype, device=x.device) if window is not None else None
    s = torch.stft(x.float() if x.dtype == torch.bfloat16 else x, fft_size, hop_size, win_length, w, center=center,
                   normalized=normalized, onesided=onesided, return_complex=True).transpose(1, 2)
    amp = torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=eps))
    fmin = 0 if fmin is None else fmin
    fmax = fs / 2 if fmax is None else fmax
    melmat = torch.from_numpy(slaney_mel_basis(fs, fft_size, num_mels, fmin, fmax).T).to(device=x.device, dtype=x.dtype)
    mel = torch.clamp(torch.matmul(amp, melmat), min=eps)
    if log_base is None:
        out = torch.log(mel)
    elif log_base == 2.0:
        out = torch.log2(mel)
    elif log_base == 10.0:
        out = torch.log10(mel)
    else:
        raise ValueError(f"log_base: {log_base} is not supported.")
    return out.transpose(1, 2)


def mel_loss(y_hat, y, **mel_params):
    """MelSpectrogramLoss.forward (losses/mel_loss.py:151-166)."""
    return F.l1_loss(mel_spectrogram(y_hat, **mel_params), mel_spectrogram(y, **mel_params))


def generator_adv_loss(outputs, average_by_discriminators=False):
    """GeneratorAdversarialLoss.forward, mse (losses/adversarial_loss.py:29-55)."""
    loss = 0.0
    for o in outputs:
        o = o[-1] if isinstance(o, (tuple, list)) else o
        loss = loss + F.mse_loss(o, torch.ones_like(o))
    return loss / len(outputs) if average_by_discriminators else loss


def discriminator_adv_loss(outputs_hat, outputs, average_by_discriminators=False):
    """DiscriminatorAdversarialLoss.forward, mse (losses/adversarial_loss.py:80-117)."""
    real, fake = 0.0, 0.0
    for oh, o in zip(outputs_hat, outputs):
        if isinstance(oh, (tuple, list)):
            oh, o = oh[-1], o[-1]
        real = real + F.mse_loss(o, torch.ones_like(o))
        fake = fake + F.mse_loss(oh, torch.zeros_like(oh))
    if average_by_discriminators:
        real, fake = real / len(outputs), fake / len(outputs)
    return real, fake


def feat_match_loss(feats_hat, feats, average_by_layers=False, average_by_discriminators=False,
                    include_final_outputs=False):
    """FeatureMatchLoss.forward (losses/feat_match_loss.py:27-54)."""
    total = 0.0
    for fh, f in zip(feats_hat, feats):
        if not include_final_outputs:
            fh, f = fh[:-1], f[:-1]
        li = 0.0
        for a, b in zip(fh, f):
            li = li + F.l1_loss(a, b.detach())
        if average_by_layers:
            li = li / len(fh)
        total = total + li
    return total / len(feats) if average_by_discriminators else total


# --------------------------------------------------------------------------- #
# optimiser + train step                                                      #
# --------------------------------------------------------------------------- #

class AdamState:
    """torch.optim.Adam(lr, betas, eps=1e-8, weight_decay=0) restated
    (bin/train.py:1750-1769; egs/ema/voc1/conf/e2w_hifigan.yaml:142-158), with the
    MultiStepLR(gamma, milestones) schedule stepped every iteration (:379-383)."""

    def __init__(self, sd, lr=1e-4, betas=(0.5, 0.9), eps=1e-8, gamma=0.5,
                 milestones=(80000, 160000, 240000, 320000)):
        self.lr0, self.betas, self.eps = lr, betas, eps
        self.gamma, self.milestones = gamma, tuple(milestones)
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}

    def lr(self):
        return self.lr0 * self.gamma ** sum(1 for m in self.milestones if self.t >= m)

    def step(self, sd, grads):
        lr = self.lr()
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        with torch.no_grad():
            for k, g in grads.items():
                if g is None:
                    continue
                self.m[k].mul_(b1).add_(g, alpha=1 - b1)
                self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
                sd[k].addcdiv_(self.m[k], denom, value=-lr / bc1)


def train_step(gsd, dsd, gparams, dparams, gopt, dopt, batch, steps, use_stft_loss=False,
               use_mel_loss=True, mel_params=None, stft_params=None, lambda_aux=45.0,
               lambda_adv=1.0, lambda_feat_match=2.0, generator_train_start_steps=1,
               discriminator_train_start_steps=0):
    """Trainer._train_step (bin/train.py:241-440) for the a2w / use_ar path.

    ``batch`` = dict(x (B,C,T'), y (B,1,T), ar (B,1,512)).  Updates ``gsd``/``dsd``
    in place, returns the dict of logged scalars (python floats).
    """
    mel_params = E2W_MEL_LOSS_PARAMS if mel_params is None else mel_params
    stft_params = DEFAULT_STFT_LOSS_PARAMS if stft_params is None else stft_params
    x, y, ar = batch["x"], batch["y"], batch["ar"]
    use_ar = _gen_defaults(gparams)["use_ar"]
    logs = {}
    if steps > generator_train_start_steps:                                   # :268
        gleaf = {k: v.detach().clone().requires_grad_(True) for k, v in gsd.items()}
        y_ = generator_forward(gleaf, gparams, x, ar)
        gen_loss = 0.0
        if use_stft_loss:                                                     # :289-297
            sc, mag = mr_stft_loss(y_, y, **stft_params)
            gen_loss = gen_loss + sc + mag
            logs["train/spectral_convergence_loss"] = float(sc)
            logs["train/log_stft_magnitude_loss"] = float(mag)
        if use_mel_loss:                                                      # :313-316
            ml = mel_loss(y_, y, **mel_params)
            gen_loss = gen_loss + ml
            logs["train/mel_loss"] = float(ml)
        gen_loss = gen_loss * lambda_aux                                      # :325
        disc_y = torch.cat([ar, y], dim=2) if use_ar else y                   # :345-349
        disc_y_ = torch.cat([ar, y_], dim=2) if use_ar else y_
        if steps > discriminator_train_start_steps:                           # :350
            p_ = discriminator_forward(dsd, dparams, disc_y_)
            adv = generator_adv_loss(p_)
            logs["train/adversarial_loss"] = float(adv)
            with torch.no_grad():
                p = discriminator_forward(dsd, dparams, disc_y)
            fm = feat_match_loss(p_, p)
            logs["train/feature_matching_loss"] = float(fm)
            adv = adv + lambda_feat_match * fm
            gen_loss = gen_loss + lambda_adv * adv
        logs["train/generator_loss"] = float(gen_loss)
        keys = list(gleaf.keys())
        grads = torch.autograd.grad(gen_loss, [gleaf[k] for k in keys], allow_unused=True)
        gopt.step(gsd, dict(zip(keys, grads)))                                # :372-383
    if steps > discriminator_train_start_steps:                               # :388
        with torch.no_grad():
            y_ = generator_forward(gsd, gparams, x, ar)                       # :390-400 (updated G)
        disc_y = torch.cat([ar, y], dim=2) if use_ar else y
        disc_y_ = torch.cat([ar, y_], dim=2) if use_ar else y_
        dleaf = {k: v.detach().clone().requires_grad_(True) for k, v in dsd.items()}
        p = discriminator_forward(dleaf, dparams, disc_y)
        p_ = discriminator_forward(dleaf, dparams, disc_y_.detach())
        real, fake = discriminator_adv_loss(p_, p)
        dis_loss = real + fake
        logs["train/real_loss"] = float(real)
        logs["train/fake_loss"] = float(fake)
        logs["train/discriminator_loss"] = float(dis_loss)
        keys = list(dleaf.keys())
        grads = torch.autograd.grad(dis_loss, [dleaf[k] for k in keys], allow_unused=True)
        dopt.step(dsd, dict(zip(keys, grads)))                                # :424-435
    return logs


# --------------------------------------------------------------------------- #
# integer indexing: collater window + chunked AR decode                       #
# --------------------------------------------------------------------------- #

def collate_window_indices(audio_len, art_len, hop_size, batch_max_steps, ar_len, start):
    """Index arithmetic of SpeechCollater.__call__ random_window branch
    (bin/train.py:983-1027, AR slice :1082-1097) for a given random ``start`` frame.
    Returns dict(art=(lo,hi), wav=(lo,hi), ar=(lo,hi), ar_left_zero_pad=n) or None if
    the item is dropped (len(art) - frames <= 0)."""
    frames = batch_max_steps // hop_size
    art_len = min(art_len, int(audio_len / hop_size))            # art[: int(len(audio)/hop)]
    if art_len - frames <= 0:
        return None
    assert 0 <= start < art_len - frames
    w0 = start * hop_size
    lo = w0 - ar_len
    return dict(art=(start, start + frames), wav=(w0, w0 + batch_max_steps),
                ar=(max(lo, 0), w0), ar_left_zero_pad=max(-lo, 0))


def ar_chunk_plan(n_frames, batch_max_steps, hop_size):
    """Chunk list of ar_loop (bin/decode.py:45-56): [(frame_lo, frame_hi, sample_lo, sample_hi)]."""
    chunk = int(batch_max_steps / hop_size)
    return [(i, min(i + chunk, n_frames), i * hop_size, min(i + chunk, n_frames) * hop_size)
            for i in range(0, n_frames, chunk)]


def ar_loop(gsd, gparams, x, batch_max_steps, hop_size):
    """ar_loop (bin/decode.py:31-83), a2w branch without WSOLA.  ``x`` is (T', C);
    returns the (T'*hop,) waveform.  Batched variant: x (B, T', C) → (B, T'*hop)."""
    p = _gen_defaults(gparams)
    past = p["ar_input"]
    batched = x.dim() == 3
    xb = x if batched else x.unsqueeze(0)
    prev = torch.zeros((xb.shape[0], p["out_channels"], past), dtype=x.dtype)
    outs = []
    for lo, hi, _, _ in ar_chunk_plan(xb.shape[1], batch_max_steps, hop_size):
        cin = xb[:, lo:hi].permute(0, 2, 1)
        cout = generator_forward(gsd, gparams, cin, prev)
        outs.append(cout[:, 0])
        if past <= batch_max_steps:                                           # :77-78
            prev = cout[:, :, -past:]
            if prev.shape[2] < past:  # short last chunk (never reused by the reference)
                prev = F.pad(prev, (past - prev.shape[2], 0))
        else:                                                                 # :79-81
            n = cout.shape[2]
            prev = torch.cat([prev[:, :, n:], cout], dim=2)
    out = torch.cat(outs, dim=1)
    return out if batched else out[0]


# --------------------------------------------------------------------------- #
# synthetic workload                                #
# --------------------------------------------------------------------------- #

def synthetic_batch(batch_size=16, in_feats=13, frames=100, hop=80, ar_len=512, seed=1234,
                    dtype=torch.float32):
    """x ~ N(0,1) (B,13,100); per item an (ar_len + frames*hop)-sample signal
    0.5 sin(2π f0 n/16000 + φ) + 0.05 N(0,1), f0 ~ U(80,300), clipped to [-1,1];
    ar = s[:ar_len], y = s[ar_len:] (mirrors the collater, bin/train.py:1082-1097)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch_size, in_feats, frames, generator=g, dtype=torch.float64)
    n = ar_len + frames * hop
    f0 = 80 + 220 * torch.rand(batch_size, 1, generator=g, dtype=torch.float64)
    phi = 2 * math.pi * torch.rand(batch_size, 1, generator=g, dtype=torch.float64)
    t = torch.arange(n, dtype=torch.float64)[None]
    s = 0.5 * torch.sin(2 * math.pi * f0 * t / 16000.0 + phi)
    s = (s + 0.05 * torch.randn(batch_size, n, generator=g, dtype=torch.float64)).clamp(-1, 1)
    return dict(x=x.to(dtype), y=s[:, None, ar_len:].to(dtype).contiguous(),
                ar=s[:, None, :ar_len].to(dtype).contiguous())
