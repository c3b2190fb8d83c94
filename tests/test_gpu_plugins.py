"""GPU parity of the remaining plugin surface: the stand-alone discriminator classes, the causal
conv layers, chunked AR decoding (ar_loop / BatchedARDecoder), load_model and the train CLI."""
import copy
import os
import warnings

import pytest
import torch
import torch.nn.functional as F
import yaml

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _small_disc_params(golden):
    dp = golden["discriminator_params"]
    return copy.deepcopy(dp["scale_discriminator_params"]), copy.deepcopy(dp["period_discriminator_params"])


def test_standalone_discriminators_vs_oracle(golden):
    from articulatory_b200 import models as M
    from oracle import torch_oracle as O
    sp, pp = _small_disc_params(golden)
    x = torch.cat([golden["batch"]["ar"], golden["batch"]["y"]], dim=2)       # (2, 1, 2512)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(3)
        pd = M.HiFiGANPeriodDiscriminator(period=5, **pp)
        mpd = M.HiFiGANMultiPeriodDiscriminator(periods=[2, 3], discriminator_params=pp)
        sd_ = M.HiFiGANScaleDiscriminator(**sp)
        msd = M.HiFiGANMultiScaleDiscriminator(scales=2, discriminator_params=sp,
                                               downsample_pooling_params={"kernel_size": 4, "stride": 2, "padding": 2})
    with torch.no_grad():
        # single period discriminator
        ref = O.period_discriminator_forward({"d." + k: v for k, v in pd.state_dict().items()}, "d", pp, 5, x)
        got = pd.to(DEV)(x.to(DEV))
        assert len(got) == len(ref) == 6
        for a, r in zip(got, ref):
            assert a.shape == r.shape and rel_err(a.float().cpu(), r) < 1e-4
        # multi period
        got = mpd.to(DEV)(x.to(DEV))
        sdm = mpd.state_dict()
        for i, p in enumerate((2, 3)):
            ref = O.period_discriminator_forward({k: v.cpu() for k, v in sdm.items()}, f"discriminators.{i}", pp, p, x)
            for a, r in zip(got[i], ref):
                assert a.shape == r.shape and rel_err(a.float().cpu(), r) < 1e-4
        # single scale
        ref = O.scale_discriminator_forward({"d." + k: v for k, v in sd_.state_dict().items()}, "d", sp, x)
        got = sd_.to(DEV)(x.to(DEV))
        assert len(got) == len(ref) == 8
        for a, r in zip(got, ref):
            assert a.shape == r.shape and rel_err(a.float().cpu(), r) < 1e-4
        # multi scale: scale 1 sees the AvgPool1d(4, 2, 2) of the input
        got = msd.to(DEV)(x.to(DEV))
        sdm = {k: v.cpu() for k, v in msd.state_dict().items()}
        xs = x
        for i in range(2):
            ref = O.scale_discriminator_forward(sdm, f"discriminators.{i}", sp, xs)
            for a, r in zip(got[i], ref):
                assert a.shape == r.shape and rel_err(a.float().cpu(), r) < 1e-4
            xs = F.avg_pool1d(xs, 4, 2, 2)


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_causal_convs_vs_torch(prec, tol):
    """CausalConv1d / CausalConvTranspose1d == reference layers/causal_conv.py semantics (pad-left + trim)."""
    from articulatory_b200.layers import CausalConv1d, CausalConvTranspose1d
    torch.manual_seed(0)
    for k, d in ((3, 1), (5, 2), (7, 3)):
        m = CausalConv1d(32, 64, k, dilation=d)
        m.precision = prec
        x = torch.randn(2, 32, 101, requires_grad=True)
        ref = F.conv1d(F.pad(x, ((k - 1) * d, 0)), m.conv.weight, m.conv.bias, dilation=d)[:, :, :101]
        dy = torch.randn_like(ref)
        gx, gw, gb = torch.autograd.grad(ref, [x, m.conv.weight, m.conv.bias], dy)
        m = m.to(DEV)
        xd = x.detach().to(DEV).requires_grad_(True)
        y = m(xd)
        assert y.shape == ref.shape and rel_err(y.cpu(), ref) < tol
        y.backward(dy.to(DEV))
        assert rel_err(xd.grad.cpu(), gx) < tol
        assert rel_err(m.conv.weight.grad.cpu(), gw) < tol and rel_err(m.conv.bias.grad.cpu(), gb) < tol
        # causality: output sample t must not depend on inputs after t
        x2 = x.detach().clone()
        x2[:, :, 60:] = 0.0
        y2 = m(x2.to(DEV))
        assert torch.equal(y2[:, :, :60].cpu(), m(x.detach().to(DEV))[:, :, :60].cpu())
    for k, s in ((8, 4), (4, 2), (16, 8)):
        m = CausalConvTranspose1d(32, 16, k, s)
        m.precision = prec
        x = torch.randn(2, 32, 50, requires_grad=True)
        ref = F.conv_transpose1d(x, m.deconv.weight, m.deconv.bias, stride=s)[:, :, :-s]
        dy = torch.randn_like(ref)
        gx, gw, gb = torch.autograd.grad(ref, [x, m.deconv.weight, m.deconv.bias], dy)
        m = m.to(DEV)
        xd = x.detach().to(DEV).requires_grad_(True)
        y = m(xd)
        assert y.shape == ref.shape and rel_err(y.cpu(), ref) < tol
        y.backward(dy.to(DEV))
        assert rel_err(xd.grad.cpu(), gx) < tol
        assert rel_err(m.deconv.weight.grad.cpu(), gw) < tol and rel_err(m.deconv.bias.grad.cpu(), gb) < tol


def _car_model(golden, precision="fp32"):
    from articulatory_b200 import models as M
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**golden["generator_params"], precision=precision)
    G.load_state_dict(golden["gsd"])
    return G


def test_ar_loop_matches_reference_fixture(golden):
    """bin/decode.py ar_loop on the CUDA generator == the reference's ar_loop output (57 frames -> 4560
    samples, last chunk short), with and without weight norm removed."""
    from articulatory_b200.bin.decode import ar_loop
    a = golden["ar_loop"]
    cfg = {"generator_params": golden["generator_params"], "batch_max_steps": a["batch_max_steps"],
           "hop_size": a["hop_size"], "dataset_mode": "a2w"}
    G = _car_model(golden).to(DEV).eval()
    with torch.no_grad():
        wav = ar_loop(G, a["art"].to(DEV), cfg)
        assert wav.shape == a["wav"].shape
        assert rel_err(wav.cpu(), a["wav"]) < 1e-4
        G.remove_weight_norm()
        assert "input_conv.weight" in G.state_dict() and "input_conv.weight_g" not in G.state_dict()
        wav2 = ar_loop(G, a["art"].to(DEV), cfg)
        assert rel_err(wav2.cpu(), a["wav"]) < 1e-4


@pytest.mark.parametrize("use_graph", [False, True])
def test_batched_decoder_equals_per_utterance_loop(golden, use_graph):
    from articulatory_b200.bin.decode import ar_loop
    from articulatory_b200.decode import BatchedARDecoder
    a = golden["ar_loop"]
    cfg = {"generator_params": golden["generator_params"], "batch_max_steps": a["batch_max_steps"],
           "hop_size": a["hop_size"], "dataset_mode": "a2w"}
    G = _car_model(golden).to(DEV).eval()
    G.remove_weight_norm()
    g = torch.Generator().manual_seed(5)
    feats = [a["art"], torch.randn(75, a["art"].shape[1], generator=g), torch.randn(25, a["art"].shape[1], generator=g),
             torch.randn(3, a["art"].shape[1], generator=g), a["art"].flip(0)]
    dec = BatchedARDecoder(G, cfg, use_graph=use_graph)
    outs = dec.decode(feats)
    assert rel_err(outs[0].cpu(), a["wav"]) < 1e-4                           # reference fixture
    with torch.no_grad():
        for f, o in zip(feats, outs):
            want = ar_loop(G, f.to(DEV), cfg)
            assert o.shape == want.shape == (len(f) * a["hop_size"],)
            assert rel_err(o.cpu(), want.cpu()) < 1e-5
    # a second call replays the captured graphs
    outs2 = dec.decode(feats)
    for o, o2 in zip(outs, outs2):
        assert torch.equal(o, o2)


def test_load_model_and_train_cli(golden, tmp_path):
    """Checkpoint written by the train entry point loads through utils.load_model (reference key names)."""
    from articulatory_b200.bin import train as train_cli
    from articulatory_b200.utils import load_model
    from tests.test_gpu_models import _train_config
    cfg = _train_config(golden, stft=True)
    cfg.update(generator_type="HiFiGANGenerator", discriminator_type="HiFiGANMultiScaleMultiPeriodDiscriminator",
               generator_params=golden["generator_params"], discriminator_params=golden["discriminator_params"],
               batch_max_steps=2000, hop_size=80, batch_size=2, sampling_rate=16000, format="npy",
               log_interval_steps=2, save_interval_steps=1000, train_max_steps=4, dataset_mode="a2w")
    conf = tmp_path / "conf.yaml"
    conf.write_text(yaml.dump(cfg))
    out = tmp_path / "exp"
    train_cli.main(["--config", str(conf), "--outdir", str(out), "--synthetic", "6", "--precision", "fp32"])
    ckpt = out / "checkpoint-4steps.pkl"
    assert ckpt.exists() and (out / "config.yml").exists()
    sd = torch.load(ckpt, map_location="cpu", weights_only=False)
    assert sd["steps"] == 4 and set(sd["model"]["generator"].keys()) == set(golden["gsd"].keys())
    assert set(sd["model"]["discriminator"].keys()) == set(golden["dsd"].keys())
    model = load_model(str(ckpt))                       # config.yml discovered beside the checkpoint
    model.remove_weight_norm()
    y = model.eval().to(DEV)(golden["batch"]["x"].to(DEV), ar=golden["batch"]["ar"].to(DEV))
    assert y.shape == (2, 1, 2000) and torch.isfinite(y).all()


@pytest.mark.parametrize("name,frames", [("e2w_hifigan.yaml", 100), ("e2w_hifigan_car.yaml", 25)])
def test_shipped_yaml_trains_and_decodes_unchanged(name, frames, tmp_path):
    """The reference's shipped recipe configs (egs/ema/voc1/conf/*.yaml, vendored byte-for-byte under tests/golden/conf)
    go through the train and decode entry points UNCHANGED: full width, the yaml's own batch size (32 / 64), losses
    (use_stft_loss: false), optimizers and schedulers; only the data (--synthetic) and the run length (--max-steps) come
    from the command line.  The default precision is the parity-gated tensor-core mode."""
    import json
    import wave

    import numpy as np
    from articulatory_b200.bin import decode as decode_cli
    from articulatory_b200.bin import train as train_cli
    conf = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conf", name)
    ycfg = yaml.load(open(conf), Loader=yaml.Loader)
    out = tmp_path / "exp"
    train_cli.main(["--config", conf, "--outdir", str(out), "--synthetic", str(ycfg["batch_size"] + 3), "--max-steps", "3"])
    ckpt = out / "checkpoint-3steps.pkl"
    assert ckpt.exists()
    sd = torch.load(ckpt, map_location="cpu", weights_only=False)
    assert sd["steps"] == 3 and "state" in sd["optimizer"]["generator"] and "param_groups" in sd["optimizer"]["discriminator"]
    # schedule gates (yaml :174-175, bin/train.py:268,388): G updates when steps > 1, D when steps > 0 -> 1 and 2 updates
    assert sd["scheduler"]["generator"]["last_epoch"] == 1
    assert sd["scheduler"]["discriminator"]["last_epoch"] == 2
    # the closing evaluation pass wrote its scalars and the intermediate results
    rows = [json.loads(l) for l in open(out / "scalars.jsonl")] if (out / "scalars.jsonl").exists() else []
    assert (rows and "eval/mel_loss" in rows[-1] and np.isfinite(rows[-1]["eval/mel_loss"])) or (out / "events").exists() or \
        any(f.startswith("events.out") for f in os.listdir(out))
    pred = out / "predictions" / "3steps"
    assert (pred / "1_gen.wav").exists() and (pred / "1_ref.wav").exists() and (pred / "4_gen.wav").exists()
    with wave.open(str(pred / "1_gen.wav")) as w:
        assert w.getnframes() == ycfg["batch_max_steps"] and w.getframerate() == ycfg["sampling_rate"]
    # decode two utterances from a kaldi-style scp of npy features with the written checkpoint + config.yml
    feats = tmp_path / "feats"
    feats.mkdir()
    rng = np.random.RandomState(0)
    lens = {"utt_a": 2 * frames + 7, "utt_b": frames - 3}
    with open(tmp_path / "feats.scp", "w") as f:
        for utt, n in lens.items():
            np.save(feats / f"{utt}.npy", rng.randn(n, 13).astype(np.float32))
            f.write(f"{utt} {feats / (utt + '.npy')}\n")
    wavs = tmp_path / "wav"
    decode_cli.main(["--feats-scp", str(tmp_path / "feats.scp"), "--checkpoint", str(ckpt), "--outdir", str(wavs)])
    for utt, n in lens.items():
        with wave.open(str(wavs / f"{utt}_gen.wav")) as w:
            assert w.getnframes() == n * ycfg["hop_size"]
            pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        assert np.abs(pcm).max() > 0


@pytest.mark.parametrize("use_ar,aux", [(True, 0), (False, 2)])
def test_device_window_cutter_matches_collater(use_ar, aux):
    """Windows cut on the device from the HBM-resident dataset == SpeechCollater's host windows under the same seeded
    RNG, bit for bit (x, y, ar), including dropped short utterances and the left-zero-padded AR context."""
    import numpy as np
    from articulatory_b200.data import DeviceWindowCutter, SpeechCollater, synthetic_utterances
    hop, steps = 80, 2000
    items = synthetic_utterances(9, n_frames=90, n_feats=13, hop_size=hop, seed=3)
    items[2] = {"audio": items[2]["audio"][:hop * 20], "art": items[2]["art"][:20]}            # too short: dropped
    items[5] = {"audio": items[5]["audio"][:hop * 61 + 17], "art": items[5]["art"]}            # art longer than the audio
    cfg = {"generator_params": {"use_ar": use_ar, "ar_input": 512, "out_channels": 1}}
    mk = lambda seed: SpeechCollater(batch_max_steps=steps, hop_size=hop, aux_context_window=aux, config=cfg,
                                     rng=np.random.RandomState(seed))
    host, dev_c = mk(11), mk(11)
    cutter = DeviceWindowCutter(items, dev_c, DEV)
    for group in ([0, 1, 2, 3], [4, 5, 6, 7, 8], [8, 0]):
        want = host([items[i] for i in group])
        got = cutter(group)
        assert torch.equal(got["x"][0].cpu(), want["x"][0]) and torch.equal(got["y"].cpu(), want["y"])
        if use_ar:
            assert torch.equal(got["ar"].cpu(), want["ar"])
        else:
            assert "ar" not in got
    # a start near the beginning exercises the zero padding of the AR context
    if use_ar:
        class First:
            def randint(self, lo, hi):
                return lo
        dev_c.rng = host.rng = First()
        want, got = host([items[0]]), cutter([0])
        assert torch.equal(got["ar"].cpu(), want["ar"]) and float(got["ar"].abs().sum()) == 0.0


def test_train_step_without_ar_conditioning(golden):
    """A non-autoregressive HiFi-GAN config (use_ar: false — the collater then emits no 'ar'): the fused train step and
    the eval step accept ar=None, D sees the bare window (reference bin/train.py:345-349 else-branch), losses match the
    oracle's step."""
    from articulatory_b200 import models as M
    from articulatory_b200.trainer import TrainStep
    from oracle import torch_oracle as O
    from tests.test_gpu_models import _train_config
    gp = dict(golden["generator_params"], use_ar=False, in_channels=13)
    torch.manual_seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**gp)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**golden["discriminator_params"])
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    dsd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    ts = TrainStep(G.to(DEV), D.to(DEV), _train_config(golden), DEV)
    b = {k: v for k, v in golden["batch"].items() if k != "ar"}
    gopt, dopt = O.AdamState(gsd), O.AdamState(dsd)
    for step in range(4):
        ref = O.train_step(gsd, dsd, gp, golden["discriminator_params"], gopt, dopt, dict(b, ar=None), step,
                           use_stft_loss=True, use_mel_loss=True)
        ts.step(b["x"].to(DEV), b["y"].to(DEV), None, use_graph=True)
        vals = ts.last_values()
        for k, v in ref.items():
            assert abs(vals[k] - v) <= 1e-3 * abs(v), (step, k, vals[k], v)
    ev = ts.eval_step(b["x"].to(DEV), b["y"].to(DEV), None)
    assert all(v == v for v in ev.values())


def test_mri_recipe_shapes_vs_oracle():
    """egs/mri/voc1/conf/mri2w_hifigan_car.yaml (vendored): 358-dim MRI features, upsampling [8, 5, 3, 2] (hop 240, 20 kHz),
    k = 16 / 10 / 6 / 4 transposed convs, CAR conditioning — generator forward on the tensor cores against the oracle."""
    from articulatory_b200 import _lib
    from articulatory_b200 import models as M
    from oracle import torch_oracle as O
    conf = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conf", "mri2w_hifigan_car.yaml")
    gp = yaml.load(open(conf), Loader=yaml.Loader)["generator_params"]
    torch.manual_seed(2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**gp, precision="bf16x3")
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    G = G.to(DEV)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, gp["in_channels"] - gp["ar_output"], 21, generator=g)
    ar = torch.rand(2, 1, gp["ar_input"], generator=g) * 2 - 1
    _lib.path_counts(reset=True)
    with torch.no_grad():
        y = G(x.to(DEV), ar=ar.to(DEV))
    pc = _lib.path_counts()
    ref = O.generator_forward(gsd, {k: v for k, v in gp.items() if k not in ("final_scale", "extra_art")}, x, ar)
    assert y.shape == ref.shape == (2, 1, 21 * 240)
    assert rel_err(y.cpu(), ref) < 1e-3
    assert pc["conv_tc_x3"] >= 77 and pc["conv_generic"] == 0, pc


def test_inference_and_register_stats(golden, tmp_path):
    """HiFiGANGenerator.inference / register_stats (reference models/hifigan.py:280-314) on a non-AR model:
    (T', C) features, normalised with the registered stats, -> (T, 1) waveform."""
    import numpy as np
    from articulatory_b200 import models as M
    from oracle import torch_oracle as O
    gp = dict(golden["generator_params"], use_ar=False, in_channels=13)
    torch.manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**gp)
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    rng = np.random.RandomState(0)
    stats = np.stack([rng.randn(13), rng.rand(13) + 0.5]).astype(np.float32)
    np.save(tmp_path / "stats.npy", stats)
    G.register_stats(str(tmp_path / "stats.npy"))
    G = G.to(DEV).eval()
    c = torch.randn(37, 13)
    y = G.inference(c.to(DEV), normalize_before=True)
    assert y.shape == (37 * 80, 1)
    cn = (c - torch.from_numpy(stats[0])) / torch.from_numpy(stats[1])
    ref = O.generator_forward(gsd, gp, cn.T[None])
    assert rel_err(y.cpu().T[None], ref) < 1e-4
    # numpy input path (reference accepts ndarray)
    y2 = G.inference(c.numpy(), normalize_before=True)
    assert torch.equal(y2, y)
    with pytest.raises(ValueError):
        _car_model(golden).to(DEV)(golden["batch"]["x"].to(DEV))      # use_ar model without `ar`


def test_off_path_options_fail_loudly():
    from articulatory_b200 import losses as L
    from articulatory_b200 import models as M
    with pytest.raises(NotImplementedError):
        L.GeneratorAdversarialLoss(loss_type="hinge")
    with pytest.raises(NotImplementedError):
        M.HiFiGANGenerator(use_spk_id=True, num_spk=4)
    with pytest.raises(NotImplementedError):
        M.HiFiGANPeriodDiscriminator(use_weight_norm=False, use_spectral_norm=True)


def test_discriminator_preamble_one_launch_equals_the_single_kernels(golden, monkeypatch):
    """DiscriminatorEngine: the one-launch input preamble (artic_disc_prep: signal assembly, AvgPool pyramid, reflect-padded
    period views) against the single-purpose kernels that take over when a row does not fit its shared memory (long
    signals) — every output of the full discriminator bit-identical, with the signal given and assembled from parts."""
    from articulatory_b200 import engine
    from articulatory_b200 import models as M
    torch.manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**copy.deepcopy(golden["discriminator_params"])).to(DEV)
    ar, y = golden["batch"]["ar"].to(DEV), golden["batch"]["y"].to(DEV)
    x = torch.cat([ar, y], dim=2)                       # (2, 1, 2512): 2512 % 3, % 5, % 7, % 11 != 0 -> reflect pads
    eng = D._ensure_ready()
    outs = {}
    for name, limit in (("one launch", 200 * 1024), ("single kernels", 0)):
        monkeypatch.setattr(engine, "_DISC_PREP_MAX_BYTES", limit)
        with torch.no_grad():
            a, _ = eng.forward(x, save=False)
            b, _ = eng.forward(None, save=False, parts=(ar.contiguous(), (y.contiguous(),)))
        torch.cuda.synchronize()
        outs[name] = [[o.t.clone() for o in lst] for lst in a], [[o.t.clone() for o in lst] for lst in b]
    for i in range(2):
        for la, lb in zip(outs["one launch"][i], outs["single kernels"][i]):
            for ta, tb in zip(la, lb):
                assert torch.equal(ta, tb)
    for la, lb in zip(outs["one launch"][0], outs["one launch"][1]):      # given signal == assembled signal
        for ta, tb in zip(la, lb):
            assert torch.equal(ta, tb)
