"""Workload constants of the shipped recipe ``egs/ema/voc1/conf/e2w_hifigan.yaml`` (reference lines cited per block),
restated here so that the benchmark and the smoke run need neither the recipe tree nor ``oracle/``.
``tests/test_host_logic.py::test_configs_match_shipped_yaml`` checks every value against the vendored yaml
(``tests/golden/conf/e2w_hifigan.yaml``).  ``synthetic_batch`` is SURVEY.md §8(d) "Config 2": the synthetic
MNGU0-shaped train batch (same arithmetic and seeds as ``oracle.torch_oracle.synthetic_batch``; equality is a test).
"""
import math

import torch

#: e2w_hifigan.yaml:33-56
E2W_GENERATOR_PARAMS = dict(
    in_channels=141, out_channels=1, channels=512, kernel_size=7,
    upsample_scales=[5, 4, 2, 2], upsample_kernel_sizes=[10, 8, 4, 4],
    resblock_kernel_sizes=[3, 7, 11], resblock_dilations=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    use_additional_convs=True, bias=True, nonlinear_activation="LeakyReLU",
    nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True,
    use_ar=True, ar_input=512, ar_hidden=256, ar_output=128)

#: e2w_hifigan.yaml:61-95
E2W_DISCRIMINATOR_PARAMS = dict(
    scales=3, scale_downsample_pooling="AvgPool1d",
    scale_downsample_pooling_params=dict(kernel_size=4, stride=2, padding=2),
    scale_discriminator_params=dict(
        in_channels=1, out_channels=1, kernel_sizes=[15, 41, 5, 3], channels=128, max_downsample_channels=1024,
        max_groups=16, bias=True, downsample_scales=[4, 4, 4, 4, 1], nonlinear_activation="LeakyReLU",
        nonlinear_activation_params={"negative_slope": 0.1}),
    follow_official_norm=True, periods=[2, 3, 5, 7, 11],
    period_discriminator_params=dict(
        in_channels=1, out_channels=1, kernel_sizes=[5, 3], channels=32, downsample_scales=[3, 3, 3, 3, 1],
        max_downsample_channels=1024, bias=True, nonlinear_activation="LeakyReLU",
        nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True, use_spectral_norm=False))

#: e2w_hifigan.yaml:102-111
E2W_MEL_LOSS_PARAMS = dict(fs=16000, fft_size=1024, hop_size=80, win_length=None, window="hann", num_mels=80,
                           fmin=0, fmax=11025, log_base=None)

#: losses/stft_loss.py:124-129 (defaults of MultiResolutionSTFTLoss; the yaml has no stft_loss_params)
DEFAULT_STFT_LOSS_PARAMS = dict(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
                                window="hann_window")


def e2w_train_config(use_stft_loss=True):
    """The train-step keys of e2w_hifigan.yaml:97-175 (``use_stft_loss: true`` is BASELINE's metric; the yaml ships
    ``false``)."""
    sched = dict(gamma=0.5, milestones=[80000, 160000, 240000, 320000])
    opt = dict(lr=1e-4, betas=[0.5, 0.9], weight_decay=0.0)
    return dict(
        use_stft_loss=use_stft_loss, use_mel_loss=True, mel_loss_params=dict(E2W_MEL_LOSS_PARAMS),
        stft_loss_params=dict(DEFAULT_STFT_LOSS_PARAMS), lambda_aux=45.0, lambda_adv=1.0, lambda_feat_match=2.0,
        use_feat_match_loss=True,
        feat_match_loss_params=dict(average_by_discriminators=False, average_by_layers=False, include_final_outputs=False),
        generator_adv_loss_params=dict(average_by_discriminators=False),
        discriminator_adv_loss_params=dict(average_by_discriminators=False),
        generator_optimizer_type="Adam", discriminator_optimizer_type="Adam",
        generator_optimizer_params=dict(opt), discriminator_optimizer_params=dict(opt),
        generator_scheduler_type="MultiStepLR", discriminator_scheduler_type="MultiStepLR",
        generator_scheduler_params=dict(sched), discriminator_scheduler_params=dict(sched),
        generator_train_start_steps=1, discriminator_train_start_steps=0,
        generator_grad_norm=-1, discriminator_grad_norm=-1)


def synthetic_batch(batch_size=16, in_feats=13, frames=100, hop=80, ar_len=512, seed=1234, dtype=torch.float32):
    """x ~ N(0,1) (B, in_feats, frames); per item an (ar_len + frames*hop)-sample signal 0.5 sin(2 pi f0 n / 16000 + phi)
    + 0.05 N(0,1), f0 ~ U(80, 300), clipped to [-1, 1]; ar = s[:ar_len], y = s[ar_len:] (the collater's slices,
    reference bin/train.py:1082-1097)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch_size, in_feats, frames, generator=g, dtype=torch.float64)
    n = ar_len + frames * hop
    f0 = 80 + 220 * torch.rand(batch_size, 1, generator=g, dtype=torch.float64)
    phi = 2 * math.pi * torch.rand(batch_size, 1, generator=g, dtype=torch.float64)
    t = torch.arange(n, dtype=torch.float64)[None]
    s = 0.5 * torch.sin(2 * math.pi * f0 * t / 16000.0 + phi)
    s = (s + 0.05 * torch.randn(batch_size, n, generator=g, dtype=torch.float64)).clamp(-1, 1)
    return dict(x=x.to(dtype), y=s[:, None, ar_len:].to(dtype).contiguous(), ar=s[:, None, :ar_len].to(dtype).contiguous())
