"""GPU parity of the speech-to-EMA inversion encoder (BiGRU, SURVEY §8 f3 / BASELINE configs[4]) against fixtures
generated from the unmodified reference (tests/golden/bigru_small.pt) and against the oracle at the shipped width."""
import os

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bigru_small.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def _build(params, sd, precision="bf16x3"):
    from articulatory_b200.models import BiGRU
    m = BiGRU(**params, precision=precision)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.eval().to(DEV)


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_bigru_forward_golden(gold, precision):
    g = gold["plain"]
    m = _build(g["params"], g["sd"], precision)
    y = m(g["x"].to(DEV))
    assert y.shape == g["y"].shape
    assert rel_err(y.cpu(), g["y"]) < 1e-4          # north_star: 1e-3 relative fp32


def test_bigru_inference_golden(gold):
    g = gold["plain"]
    m = _build(g["params"], g["sd"])
    m.register_buffer("mean", g["mean"].to(DEV))
    m.register_buffer("scale", g["scale"].to(DEV))
    y = m.inference(g["c"].to(DEV), normalize_before=True)
    assert y.shape == (41, 12)
    assert rel_err(y.cpu(), g["y_inference"]) < 1e-4


def test_bigru_ar_tanh_golden(gold):
    g = gold["ar_tanh"]
    m = _build(g["params"], g["sd"])
    y = m(g["x"].to(DEV), ar=g["ar"].to(DEV))
    assert y.shape == (2, 12, 33) and float(y.abs().max()) <= 1.0
    assert rel_err(y.cpu(), g["y"]) < 1e-4


@pytest.mark.parametrize("cin,N,T", [(80, 3, 120), (1024, 20, 257), (13, 1, 64)])
def test_bigru_full_width_vs_oracle(cin, N, T):
    """Shipped width (2 x BiGRU(256) -> 128 -> BN -> 12) with MFCC- / mel- / HuBERT-sized inputs; the 1024-wide input
    projection must run on the split-operand tensor-core kernel."""
    from articulatory_b200 import _lib
    from articulatory_b200.models import BiGRU
    from oracle import inversion_oracle as I
    torch.manual_seed(3)
    m = BiGRU(in_channels=cin, hidden_size=256, out_channels=12)
    with torch.no_grad():
        m.bn.running_mean.normal_(0, 0.2)
        m.bn.running_var.uniform_(0.5, 1.5)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.eval().to(DEV)
    x = torch.randn(N, cin, T)
    _lib.path_counts(reset=True)
    y = m(x.to(DEV))
    pc = _lib.path_counts()
    want = I.bigru_forward(sd, x)
    assert rel_err(y.cpu(), want) < 1e-4
    assert pc["conv_tc_x3"] >= (2 if cin % 16 == 0 else 1), pc     # gi1 (when the width allows) and gi2


def test_bigru_refuses_training_mode():
    from articulatory_b200.models import BiGRU
    m = BiGRU(in_channels=13, hidden_size=32, out_channels=12).to(DEV)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 13, 8, device=DEV))
