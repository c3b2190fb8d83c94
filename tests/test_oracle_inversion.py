"""The inversion (speech -> EMA) oracle against fixtures generated from the unmodified reference BiGRU
(tests/golden/make_golden_inversion.py), and live against the reference where it is mounted."""
import os

import pytest
import torch

from oracle import inversion_oracle as I
from oracle import ref_shim
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bigru_small.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def test_bigru_forward_golden(gold):
    g = gold["plain"]
    y = I.bigru_forward(g["sd"], g["x"])
    assert y.shape == g["y"].shape == (3, 12, 57)
    assert rel_err(y, g["y"]) < 1e-5


def test_bigru_inference_golden(gold):
    g = gold["plain"]
    y = I.bigru_inference(g["sd"], g["c"], normalize_before=True, mean=g["mean"], scale=g["scale"])
    assert y.shape == (41, 12)
    assert rel_err(y, g["y_inference"]) < 1e-5


def test_bigru_ar_tanh_golden(gold):
    g = gold["ar_tanh"]
    y = I.bigru_forward(g["sd"], g["x"], ar=g["ar"], use_tanh=True)
    assert y.shape == (2, 12, 33) and float(y.abs().max()) <= 1.0
    assert rel_err(y, g["y"]) < 1e-5


def test_gru_direction_matches_time_reversal():
    """The backward direction is the forward recurrence on the time-reversed sequence, re-reversed."""
    torch.manual_seed(0)
    H, C = 5, 3
    w_ih, w_hh, b_ih, b_hh = torch.randn(3 * H, C), torch.randn(3 * H, H), torch.randn(3 * H), torch.randn(3 * H)
    x = torch.randn(2, 9, C)
    a = I.gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=True)
    b = I.gru_direction(x.flip(1), w_ih, w_hh, b_ih, b_hh).flip(1)
    assert torch.allclose(a, b, atol=1e-6)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_bigru_vs_reference_full_width():
    """Full width of the shipped default (80 -> 2 x BiGRU(256) -> 128 -> 12), random weights, eval mode."""
    ref_shim.install()
    from articulatory.models.pytorch_models import BiGRU
    torch.manual_seed(3)
    m = BiGRU(in_channels=80, hidden_size=256, out_channels=12).eval()
    with torch.no_grad():
        m.bn.running_mean.normal_(0, 0.2)
        m.bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(2, 80, 120)
    with torch.no_grad():
        want = m(x)
    got = I.bigru_forward(m.state_dict(), x)
    assert rel_err(got, want) < 1e-5


def test_plugin_class_has_the_reference_state_dict(gold):
    """articulatory_b200.models.BiGRU (parameter containers of the CUDA path) loads the reference's state_dict
    strictly, for the plain and the AR + tanh variants (no GPU needed)."""
    from articulatory_b200.models import BiGRU
    for k in ("plain", "ar_tanh"):
        m = BiGRU(**gold[k]["params"])
        res = m.load_state_dict(gold[k]["sd"], strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(NotImplementedError):
        BiGRU(use_spk_emb=True)
