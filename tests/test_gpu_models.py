"""GPU parity of the plugin modules and the fused train step against the oracle and the
reference-generated golden fixtures (tests/golden/small_e2w.pt).  `pytest -m gpu`."""
import copy
import warnings

import pytest
import torch

from tests.helpers import check_digest, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(golden, precision="fp32"):
    from articulatory_b200 import models as M
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**golden["generator_params"], precision=precision)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**golden["discriminator_params"], precision=precision)
    G.load_state_dict(golden["gsd"])
    D.load_state_dict(golden["dsd"])
    return G.to(DEV), D.to(DEV)


def test_generator_forward_golden(golden):
    G, _ = _build(golden)
    b = golden["batch"]
    with torch.no_grad():
        y = G(b["x"].to(DEV), ar=b["ar"].to(DEV))
    assert y.shape == (2, 1, 2000)
    assert rel_err(y.cpu(), golden["g_out"]) < 1e-4          # north_star: 1e-3 relative fp32


def test_generator_forward_bf16_reported(golden):
    """bf16 speed mode: error is REPORTED against the fp32 reference, gate is loose."""
    G, _ = _build(golden, "bf16")
    b = golden["batch"]
    with torch.no_grad():
        y = G(b["x"].to(DEV), ar=b["ar"].to(DEV))
    e = rel_err(y.cpu(), golden["g_out"])
    print(f"bf16 generator waveform rel err = {e:.3e}")
    assert e < 5e-2


def test_discriminator_forward_golden(golden):
    _, D = _build(golden)
    x = torch.cat([golden["batch"]["ar"], golden["g_out"]], dim=2).to(DEV)
    with torch.no_grad():
        outs = D(x)
    assert len(outs) == 8
    for i, (mine, ref) in enumerate(zip(outs, golden["d_out"])):
        assert len(mine) == len(ref)
        for j, (a, r) in enumerate(zip(mine[:-1], ref[:-1])):
            check_digest(a.float().cpu(), r, 1e-4, f"D{i} fmap{j}")
        assert mine[-1].shape == ref[-1].shape
        assert rel_err(mine[-1].cpu(), ref[-1]) < 1e-4


def test_autograd_gradients_golden(golden):
    """Same losses as tests/golden/make_golden.py 'g_grads' / 'd_grads', through the plugin
    modules + loss modules + torch autograd (the drop-in path)."""
    from articulatory_b200 import losses as L
    from oracle import torch_oracle as O
    G, D = _build(golden)
    b = {k: v.to(DEV) for k, v in golden["batch"].items()}
    mel = L.MelSpectrogramLoss(**O.E2W_MEL_LOSS_PARAMS).to(DEV)
    y_ = G(b["x"], ar=b["ar"])
    for p in D.parameters():
        p.requires_grad_(False)
    p_ = D(torch.cat([b["ar"], y_], dim=2))
    with torch.no_grad():
        p = D(torch.cat([b["ar"], b["y"]], dim=2))
    loss = 45.0 * mel(y_, b["y"]) + L.GeneratorAdversarialLoss(False)(p_) + 2.0 * L.FeatureMatchLoss(False, False, False)(p_, p)
    assert abs(loss.item() - float(golden["gen_loss"])) < 1e-4 * float(golden["gen_loss"])
    loss.backward()
    for k, v in G.named_parameters():
        check_digest(v.grad.cpu(), golden["g_grads"][k], 2e-3, f"G grad {k}")
    for p_i in D.parameters():
        p_i.requires_grad_(True)
    pr = D(torch.cat([b["ar"], b["y"]], dim=2))
    pf = D(torch.cat([b["ar"], y_.detach()], dim=2))
    r, f = L.DiscriminatorAdversarialLoss(False)(pf, pr)
    (r + f).backward()
    for k, v in D.named_parameters():
        check_digest(v.grad.cpu(), golden["d_grads"][k], 2e-3, f"D grad {k}")


def _train_config(golden, stft=True):
    from oracle import torch_oracle as O
    return dict(use_stft_loss=stft, use_mel_loss=True, mel_loss_params=O.E2W_MEL_LOSS_PARAMS,
                stft_loss_params=O.DEFAULT_STFT_LOSS_PARAMS, lambda_aux=45.0, lambda_adv=1.0, lambda_feat_match=2.0,
                use_feat_match_loss=True,
                feat_match_loss_params=dict(average_by_discriminators=False, average_by_layers=False, include_final_outputs=False),
                generator_adv_loss_params=dict(average_by_discriminators=False),
                discriminator_adv_loss_params=dict(average_by_discriminators=False),
                generator_optimizer_params=dict(lr=1e-4, betas=[0.5, 0.9], weight_decay=0.0),
                discriminator_optimizer_params=dict(lr=1e-4, betas=[0.5, 0.9], weight_decay=0.0),
                generator_scheduler_params=dict(gamma=0.5, milestones=[80000, 160000, 240000, 320000]),
                discriminator_scheduler_params=dict(gamma=0.5, milestones=[80000, 160000, 240000, 320000]),
                generator_train_start_steps=1, discriminator_train_start_steps=0,
                generator_grad_norm=-1, discriminator_grad_norm=-1)


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_steps_golden(golden, use_graph):
    """Four fused train steps == four reference Trainer._train_step calls (logged scalars)."""
    from articulatory_b200.trainer import LOG_KEYS, TrainStep
    G, D = _build(golden)
    ts = TrainStep(G, D, _train_config(golden), DEV)
    b = {k: v.to(DEV) for k, v in golden["batch"].items()}
    for step, ref_logs in enumerate(golden["train_logs"]):
        ts.step(b["x"], b["y"], b["ar"], use_graph=use_graph)
        vals = ts.last_values()
        for k, v in ref_logs.items():
            assert abs(vals[k] - v) <= 1e-3 * abs(v), (step, k, vals[k], v)
    # weight deltas: loose, see tests/test_oracle_golden.py::test_train_steps
    gsd = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    for k, d in golden["gsd_delta"].items():
        check_digest(gsd[k] - golden["gsd"][k], d, 0.9, f"G delta {k}", sum_rtol=0.3)
    dsd = {k: v.detach().cpu() for k, v in D.state_dict().items()}
    for k, d in golden["dsd_delta"].items():
        check_digest(dsd[k] - golden["dsd"][k], d, 0.9, f"D delta {k}", sum_rtol=0.3)
    # ... and tight against the ORACLE run live on the same four steps: direction and size of every network's update.
    # (Element-wise the deltas are fragile — Adam's first steps are sign-like, so a weight whose gradient is ~0 flips
    # with the summation order; over a whole network the update must agree.)
    from oracle import torch_oracle as O
    ogsd = {k: v.clone() for k, v in golden["gsd"].items()}
    odsd = {k: v.clone() for k, v in golden["dsd"].items()}
    gopt, dopt = O.AdamState(ogsd), O.AdamState(odsd)
    for step in range(len(golden["train_logs"])):
        O.train_step(ogsd, odsd, golden["generator_params"], golden["discriminator_params"], gopt, dopt, golden["batch"], step,
                     use_stft_loss=True, use_mel_loss=True)
    # thresholds: the fp32 oracle against ITSELF in fp64 gives cos 0.980 / norm ratio 0.986 for G (two sign-like Adam
    # updates) and 0.9999999 / 1.000002 for D (three updates of well-conditioned gradients)
    for name, mine, ref, init, cmin, ntol in (("G", gsd, ogsd, golden["gsd"], 0.9, 0.06), ("D", dsd, odsd, golden["dsd"], 0.995, 0.02)):
        dm = torch.cat([(mine[k] - init[k]).flatten().double() for k in init])
        dr = torch.cat([(ref[k] - init[k]).flatten().double() for k in init])
        cos = float(dm @ dr / (dm.norm() * dr.norm()))
        assert cos > cmin and abs(float(dm.norm() / dr.norm()) - 1.0) < ntol, (name, cos, float(dm.norm() / dr.norm()))
    # a 5th step replays the captured graph (when enabled) and must stay finite
    ts.step(b["x"], b["y"], b["ar"], use_graph=use_graph)
    assert all(map(lambda v: v == v and abs(v) < 1e6, ts.last_values().values()))


def test_train_steps_golden_with_split_discriminator_backward(golden, monkeypatch):
    """The alternative schedule (D backward over the real half in the background of the G backward, trainer._SPLIT_DBWD)
    is the same arithmetic: the reference's logged scalars of four steps, graph mode."""
    from articulatory_b200 import trainer
    monkeypatch.setattr(trainer, "_SPLIT_DBWD", True)
    G, D = _build(golden)
    ts = trainer.TrainStep(G, D, _train_config(golden), DEV)
    b = {k: v.to(DEV) for k, v in golden["batch"].items()}
    for step, ref_logs in enumerate(golden["train_logs"]):
        ts.step(b["x"], b["y"], b["ar"], use_graph=True)
        vals = ts.last_values()
        for k, v in ref_logs.items():
            assert abs(vals[k] - v) <= 1e-3 * abs(v), (step, k, vals[k], v)


def test_eval_step_vs_oracle(golden):
    """TrainStep.eval_step == the reference Trainer._eval_step arithmetic (bin/train.py:470-603), restated with the
    oracle's functions on the golden weights / batch: nine eval/* losses, no parameter change."""
    from articulatory_b200.trainer import TrainStep
    from oracle import torch_oracle as O
    G, D = _build(golden)
    cfg = _train_config(golden)
    ts = TrainStep(G, D, cfg, DEV)
    b = golden["batch"]
    before = {k: v.detach().clone() for k, v in list(G.state_dict().items()) + list(D.state_dict().items())}
    got = ts.eval_step(b["x"].to(DEV), b["y"].to(DEV), b["ar"].to(DEV))
    gsd, dsd = golden["gsd"], golden["dsd"]
    with torch.no_grad():
        y_ = O.generator_forward(gsd, golden["generator_params"], b["x"], b["ar"])
        sc, mag = O.mr_stft_loss(y_.squeeze(1), b["y"].squeeze(1))
        mel = O.mel_loss(y_, b["y"], **O.E2W_MEL_LOSS_PARAMS)
        p_ = O.discriminator_forward(dsd, golden["discriminator_params"], torch.cat([b["ar"], y_], dim=2))
        p = O.discriminator_forward(dsd, golden["discriminator_params"], torch.cat([b["ar"], b["y"]], dim=2))
        adv = O.generator_adv_loss(p_)
        fm = O.feat_match_loss(p_, p)
        real, fake = O.discriminator_adv_loss(p_, p)
        gen = 45.0 * (sc + mag + mel) + 1.0 * adv + 1.0 * 2.0 * fm
    want = {"eval/spectral_convergence_loss": sc, "eval/log_stft_magnitude_loss": mag, "eval/mel_loss": mel,
            "eval/adversarial_loss": adv, "eval/feature_matching_loss": fm, "eval/generator_loss": gen,
            "eval/real_loss": real, "eval/fake_loss": fake, "eval/discriminator_loss": real + fake}
    for k, v in want.items():
        assert abs(got[k] - float(v)) <= 1e-3 * abs(float(v)), (k, got[k], float(v))
    after = dict(list(G.state_dict().items()) + list(D.state_dict().items()))
    assert all(torch.equal(before[k], after[k]) for k in before)
    avg = ts.read_eval_logs(1)
    assert abs(avg["eval/mel_loss"] - got["eval/mel_loss"]) < 1e-6


def test_full_width_generator_and_discriminator_vs_oracle():
    """e2w_hifigan.yaml widths, B=2: waveform and all 54 discriminator outputs vs the oracle."""
    from articulatory_b200 import models as M
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS)
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    dsd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    G, D = G.to(DEV), D.to(DEV)
    b = O.synthetic_batch(2)
    with torch.no_grad():
        y = G(b["x"].to(DEV), ar=b["ar"].to(DEV))
        y_ref = O.generator_forward(gsd, O.E2W_GENERATOR_PARAMS, b["x"], b["ar"])
        assert y.shape == (2, 1, 8000)
        assert rel_err(y.cpu(), y_ref) < 1e-3
        din = torch.cat([b["ar"], b["y"]], 2)
        outs = D(din.to(DEV))
        ref = O.discriminator_forward(dsd, O.E2W_DISCRIMINATOR_PARAMS, din)
        for lo, lr_ in zip(outs, ref):
            for a, r in zip(lo, lr_):
                assert a.shape == r.shape
                assert rel_err(a.float().cpu(), r) < 1e-3


# --------------------------------------------------------------------------------------------------------------
# The tensor-core modes at e2w_hifigan.yaml width, end to end against the oracle.  "bf16x3" is the PARITY-GATED
# tensor-core configuration (north_star: 1e-3 relative fp32 on waveforms / losses); "bf16" is the speed mode whose
# error is RECORDED (printed, written to gpurun_out/) and gated at the level the reference itself shows under bf16
# autocast (1.3e-2 waveform, up to 8e-3 on D logits).
# --------------------------------------------------------------------------------------------------------------
_FULL_GATES = {"bf16x3": dict(wave=1e-3, dout=1e-3, loss=1e-3), "bf16": dict(wave=5e-2, dout=5e-2, loss=5e-2)}


def _record(name, payload):
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"parity_{name}.json"), "w") as f:
        json.dump(payload, f, indent=1)
    print(f"[parity] {name}: {json.dumps(payload)[:2000]}")


def _full_width(precision, seed=0):
    from articulatory_b200 import models as M
    from oracle import torch_oracle as O
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**O.E2W_GENERATOR_PARAMS, precision=precision)
        D = M.HiFiGANMultiScaleMultiPeriodDiscriminator(**O.E2W_DISCRIMINATOR_PARAMS, precision=precision)
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    dsd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    return G.to(DEV), D.to(DEV), gsd, dsd


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_full_width_tensor_core_forward_vs_oracle(precision):
    """G waveform + all 54 D outputs at full width on the tcgen05 path; every eligible layer must take it."""
    from articulatory_b200 import _lib
    from oracle import torch_oracle as O
    G, D, gsd, dsd = _full_width(precision)
    gate = _FULL_GATES[precision]
    b = O.synthetic_batch(2)
    _lib.path_counts(reset=True)
    with torch.no_grad():
        y = G(b["x"].to(DEV), ar=b["ar"].to(DEV))
        pc_g = _lib.path_counts(reset=True)
        y_ref = O.generator_forward(gsd, O.E2W_GENERATOR_PARAMS, b["x"], b["ar"])
        e_wave = rel_err(y.cpu(), y_ref)
        din = torch.cat([b["ar"], b["y"]], 2)
        outs = D(din.to(DEV))
        pc_d = _lib.path_counts(reset=True)
        ref = O.discriminator_forward(dsd, O.E2W_DISCRIMINATOR_PARAMS, din)
    e_d = [[rel_err(a.float().cpu(), r) for a, r in zip(lo, lr_)] for lo, lr_ in zip(outs, ref)]
    _record(f"forward_{precision}", {"waveform": e_wave, "d_outputs_max": max(map(max, e_d)), "d_outputs": e_d,
                                      "paths_generator": pc_g, "paths_discriminator": pc_d})
    tc = "conv_tc_x3" if precision == "bf16x3" else "conv_tc"
    # generator: 1 input conv + 4 upsamples (stride phases share one launch group) + 72 MRF convs on the tensor cores; the
    # 5 AR linears, and the 32 -> 1 output conv (channel-1 kernel) are the only others
    # (bf16: the 18 residual units of the C = 32 / 64 stages run as fused two-conv launches)
    assert pc_g[tc] + 2 * pc_g["conv_tc_fused"] >= 77 and pc_g["conv_generic"] <= 5 and pc_g["conv_c1"] == 1, pc_g
    assert pc_g["conv_tc_fused"] == (18 if precision == "bf16" else 0), pc_g
    # discriminator: per chain only the first (C_in = 1) and the logits (C_out = 1) convs are channel-1 kernels
    # (the channel-1 kernels are written for the bf16 mode; with fp32 storage the first layers use the generic fp32 kernel)
    assert pc_d["conv_c1"] + pc_d["conv_generic"] == 16 and pc_d[tc] >= 37, pc_d
    if precision == "bf16":
        assert pc_d["conv_generic"] == 0, pc_d
    assert e_wave < gate["wave"], e_wave
    assert max(map(max, e_d)) < gate["dout"], e_d


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_full_width_train_steps_vs_oracle(precision):
    """Four train steps (B = 2, mel + MR-STFT + adversarial + feature matching, CUDA graph from the third on) at full
    width on the tcgen05 path against oracle.train_step: the nine logged losses of every step."""
    from articulatory_b200 import _lib
    from articulatory_b200.trainer import TrainStep
    from oracle import torch_oracle as O
    G, D, gsd, dsd = _full_width(precision)
    gate = _FULL_GATES[precision]["loss"]
    b = O.synthetic_batch(2)
    cfg = _train_config(None)
    ts = TrainStep(G, D, cfg, DEV)
    gopt, dopt = O.AdamState(gsd), O.AdamState(dsd)
    bd = {k: v.to(DEV) for k, v in b.items()}
    errs, worst = [], 0.0
    _lib.path_counts(reset=True)
    for step in range(4):
        ref = O.train_step(gsd, dsd, O.E2W_GENERATOR_PARAMS, O.E2W_DISCRIMINATOR_PARAMS, gopt, dopt, b, step,
                           use_stft_loss=True, use_mel_loss=True)
        ts.step(bd["x"], bd["y"], bd["ar"], use_graph=True)
        vals = ts.last_values()
        e = {k: abs(vals[k] - v) / max(abs(v), 1e-12) for k, v in ref.items()}
        errs.append(e)
        worst = max([worst] + list(e.values()))
    pc = _lib.path_counts()
    _record(f"train_steps_{precision}", {"worst_rel_err": worst, "per_step": errs, "paths": pc})
    tcw = "wgrad_tc_x3" if precision == "bf16x3" else "wgrad_tc"
    assert pc[tcw] > 0 and pc["wgrad_generic"] <= 3 * 5, pc      # only the AR linears may use the generic wgrad
    if precision == "bf16":
        assert pc["conv_generic"] == 0, pc                       # every conv of the bf16 step: tcgen05 or a channel-1 kernel
    assert worst < gate, errs


def test_cpu_input_fails_loudly(golden):
    from articulatory_b200._lib import ArticError
    G, _ = _build(golden)
    with pytest.raises(ArticError):
        G(golden["batch"]["x"], ar=golden["batch"]["ar"])
