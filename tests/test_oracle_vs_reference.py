"""Live check of the oracle against the unmodified reference at FULL e2w_hifigan.yaml
width.  Runs only where /root/reference exists (the build container)."""
import warnings

import pytest
import torch

from oracle import ref_shim
from oracle import torch_oracle as O
from tests.helpers import rel_err

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")]


@pytest.fixture(scope="module")
def ref():
    import yaml
    ref_shim.install()
    import articulatory.losses as L
    import articulatory.models as M
    cfg = yaml.safe_load(open(ref_shim.REF_ROOT + "/egs/ema/voc1/conf/e2w_hifigan.yaml"))
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**cfg["generator_params"])
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**cfg["discriminator_params"])
    return dict(cfg=cfg, G=G, D=D, L=L, M=M)


def test_yaml_matches_oracle_constants(ref):
    cfg = ref["cfg"]
    assert cfg["generator_params"] == O.E2W_GENERATOR_PARAMS
    assert cfg["discriminator_params"] == O.E2W_DISCRIMINATOR_PARAMS
    assert cfg["mel_loss_params"] == O.E2W_MEL_LOSS_PARAMS


def test_state_dict_names_and_param_counts(ref):
    gsd, dsd = ref["G"].state_dict(), ref["D"].state_dict()
    gi = O.init_generator_state(ref["cfg"]["generator_params"])
    di = O.init_discriminator_state(ref["cfg"]["discriminator_params"])
    assert {k: tuple(v.shape) for k, v in gsd.items()} == {k: tuple(v.shape) for k, v in gi.items()}
    assert {k: tuple(v.shape) for k, v in dsd.items()} == {k: tuple(v.shape) for k, v in di.items()}
    assert sum(v.numel() for v in gsd.values()) == 13467778
    assert sum(v.numel() for v in dsd.values()) == 70711277


def test_full_width_forward(ref):
    cfg = ref["cfg"]
    b = O.synthetic_batch(2)
    gsd = {k: v.detach() for k, v in ref["G"].state_dict().items()}
    dsd = {k: v.detach() for k, v in ref["D"].state_dict().items()}
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y_ref = ref["G"](b["x"], ar=b["ar"])
        y = O.generator_forward(gsd, cfg["generator_params"], b["x"], b["ar"])
        assert y.shape == (2, 1, 8000) and rel_err(y, y_ref) < 1e-5
        din = torch.cat([b["ar"], b["y"]], 2)
        for lr_, lo in zip(ref["D"](din), O.discriminator_forward(dsd, cfg["discriminator_params"], din)):
            for a, c in zip(lr_, lo):
                assert a.shape == c.shape and rel_err(c, a) < 1e-5
        sc_r, mag_r = ref["L"].MultiResolutionSTFTLoss()(y_ref, b["y"])
        sc, mag = O.mr_stft_loss(y_ref, b["y"])
        assert abs(float(sc) - float(sc_r)) < 1e-6 and abs(float(mag) - float(mag_r)) < 1e-5
        ml_r = ref["L"].MelSpectrogramLoss(**cfg["mel_loss_params"])(y_ref, b["y"])
        assert abs(float(O.mel_loss(y_ref, b["y"], **cfg["mel_loss_params"])) - float(ml_r)) < 1e-5


def test_car_yaml_extra_keys(ref):
    """e2w_hifigan_car.yaml carries final_scale / extra_art, which the reference class
    rejects; the oracle ignores them."""
    import yaml
    cfg = yaml.safe_load(open(ref_shim.REF_ROOT + "/egs/ema/voc1/conf/e2w_hifigan_car.yaml"))
    gp = cfg["generator_params"]
    assert "final_scale" in gp and "extra_art" in gp
    with pytest.raises(TypeError):
        ref["M"].HiFiGANGenerator(**gp)
    sd = O.init_generator_state(gp)
    b = O.synthetic_batch(1, frames=25)
    assert O.generator_forward(sd, gp, b["x"], b["ar"]).shape == (1, 1, 2000)


def test_eval_step_arithmetic_vs_reference_trainer():
    """The nine eval/* losses as TrainStep.eval_step composes them (tests/test_gpu_models.py::test_eval_step_vs_oracle
    uses the same oracle formulas) == the reference Trainer._eval_step (bin/train.py:470-603) on the golden
    reduced-width networks."""
    import os

    ref_shim.install()
    import articulatory.bin.train as ref_train
    import articulatory.losses as L
    import articulatory.models as M
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "small_e2w.pt"), map_location="cpu",
                      weights_only=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**gold["generator_params"])
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**gold["discriminator_params"])
    G.load_state_dict(gold["gsd"])
    D.load_state_dict(gold["dsd"])
    config = dict(outdir="/tmp", generator_params=gold["generator_params"], use_stft_loss=True,
                  use_subband_stft_loss=False, use_mel_loss=True, use_inter_loss=False, use_ph_loss=False,
                  lambda_aux=45.0, lambda_adv=1.0, lambda_feat_match=2.0, use_feat_match_loss=True,
                  generator_train_start_steps=1, discriminator_train_start_steps=0, generator_grad_norm=-1,
                  discriminator_grad_norm=-1, train_max_steps=100)
    criterion = {"gen_adv": L.GeneratorAdversarialLoss(average_by_discriminators=False),
                 "dis_adv": L.DiscriminatorAdversarialLoss(average_by_discriminators=False),
                 "stft": L.MultiResolutionSTFTLoss(), "mel": L.MelSpectrogramLoss(**O.E2W_MEL_LOSS_PARAMS),
                 "feat_match": L.FeatureMatchLoss(False, False, False)}
    tr = ref_train.Trainer(0, 0, None, None, {"generator": G, "discriminator": D}, criterion, {}, {}, config)
    b = gold["batch"]
    tr._eval_step({"x": (b["x"],), "y": b["y"], "ar": b["ar"]})
    want = dict(tr.total_eval_loss)
    with torch.no_grad():
        y_ = O.generator_forward(gold["gsd"], gold["generator_params"], b["x"], b["ar"])
        sc, mag = O.mr_stft_loss(y_.squeeze(1), b["y"].squeeze(1))
        mel = O.mel_loss(y_, b["y"], **O.E2W_MEL_LOSS_PARAMS)
        p_ = O.discriminator_forward(gold["dsd"], gold["discriminator_params"], torch.cat([b["ar"], y_], dim=2))
        p = O.discriminator_forward(gold["dsd"], gold["discriminator_params"], torch.cat([b["ar"], b["y"]], dim=2))
        adv, fm = O.generator_adv_loss(p_), O.feat_match_loss(p_, p)
        real, fake = O.discriminator_adv_loss(p_, p)
        gen = 45.0 * (sc + mag + mel) + 1.0 * adv + 1.0 * 2.0 * fm
    got = {"eval/spectral_convergence_loss": sc, "eval/log_stft_magnitude_loss": mag, "eval/mel_loss": mel,
           "eval/adversarial_loss": adv, "eval/feature_matching_loss": fm, "eval/generator_loss": gen,
           "eval/real_loss": real, "eval/fake_loss": fake, "eval/discriminator_loss": real + fake}
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(float(got[k]) - v) <= 1e-4 * abs(v), (k, float(got[k]), v)
