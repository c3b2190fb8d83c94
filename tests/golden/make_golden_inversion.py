"""Golden fixture of the speech-to-EMA inversion forward from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_inversion.py

Builds ``articulatory.models.BiGRU`` (models/pytorch_models.py:22) at reduced width (hidden 24, 13 MFCC-like
inputs, 12 EMA outputs; one variant with AR conditioning + tanh), gives BatchNorm non-trivial running
statistics, and records weights, inputs and eval-mode outputs in tests/golden/bigru_small.pt.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from articulatory.models.pytorch_models import BiGRU  # noqa: E402


def build(seed, **kw):
    torch.manual_seed(seed)
    m = BiGRU(**kw)
    with torch.no_grad():
        m.bn.running_mean.copy_(torch.randn(128) * 0.3)
        m.bn.running_var.copy_(torch.rand(128) + 0.5)
        m.bn.weight.copy_(torch.rand(128) + 0.5)
        m.bn.bias.copy_(torch.randn(128) * 0.1)
    return m.eval()


def main():
    out = {}
    g = torch.Generator().manual_seed(1234)
    plain = build(0, in_channels=13, hidden_size=24, out_channels=12)
    x = torch.randn(3, 13, 57, generator=g)
    with torch.no_grad():
        y = plain(x)
        c = torch.randn(41, 13, generator=g)
        mean, scale = torch.randn(13, generator=g), torch.rand(13, generator=g) + 0.5
        plain.register_buffer("mean", mean)
        plain.register_buffer("scale", scale)
        yi = plain.inference(c, normalize_before=True)
    sd = {k: v.clone() for k, v in plain.state_dict().items() if k not in ("mean", "scale")}
    out["plain"] = dict(params=dict(in_channels=13, hidden_size=24, out_channels=12), sd=sd, x=x, y=y,
                        c=c, mean=mean, scale=scale, y_inference=yi)
    arm = build(1, in_channels=13 + 16, hidden_size=24, out_channels=12, use_ar=True, ar_input=40, ar_hidden=32,
                ar_output=16, use_tanh=True)
    x2 = torch.randn(2, 13, 33, generator=g)
    ar = torch.rand(2, 1, 40, generator=g) * 2 - 1
    with torch.no_grad():
        y2 = arm(x2, ar=ar)
    out["ar_tanh"] = dict(params=dict(in_channels=29, hidden_size=24, out_channels=12, use_ar=True, ar_input=40,
                                      ar_hidden=32, ar_output=16, use_tanh=True),
                          sd={k: v.clone() for k, v in arm.state_dict().items()}, x=x2, ar=ar, y=y2)
    path = os.path.join(ROOT, "tests", "golden", "bigru_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
