"""CPU oracle of the speech-to-EMA inversion forward (BASELINE config 5).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing in the product path imports this file.

Restates ``articulatory.models.BiGRU`` (models/pytorch_models.py:22-77) in eval mode on a plain
``state_dict`` with the reference's key names: two single-layer bidirectional GRUs (hidden 256, batch first),
Linear 512 -> 128 (+ dropout: identity in eval), BatchNorm1d(128) over the feature axis with running
statistics, Linear 128 -> out (+ tanh when ``use_tanh``), optional 512-sample AR conditioning through
``PastFCEncoder`` (layers/pytorch_layers.py:426-461) repeated over time and concatenated to the input
(:57-60).  The GRU cell is written out gate by gate (torch.nn.GRU docs: r, z, n in that order in the stacked
weights) rather than calling ``torch.nn.GRU``.

Pinned against the reference itself in tests/test_oracle_vs_reference.py (where /root/reference exists) and
against tests/golden/bigru_small.pt generated from the unmodified reference by
tests/golden/make_golden_inversion.py.
"""
import torch
import torch.nn.functional as F

from oracle.torch_oracle import past_fc_encoder   # PastFCEncoder (layers/pytorch_layers.py:426-461), pinned there


def gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """One direction of a single-layer GRU, batch first.  x (N, T, C) -> (N, T, H).

        r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)
        z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn))
        h' = (1 - z) * n + z * h
    """
    N, T, _ = x.shape
    H = w_hh.shape[1]
    gi_all = x @ w_ih.t() + b_ih                         # (N, T, 3H): the input projections of every step
    h = x.new_zeros(N, H)
    out = x.new_empty(N, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gi = gi_all[:, t]
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        out[:, t] = h
    return out


def bigru_layer(sd, prefix, x):
    """Bidirectional single-layer GRU: concat(forward, backward) along features (torch.nn.GRU, bidirectional)."""
    f = gru_direction(x, sd[f"{prefix}.weight_ih_l0"], sd[f"{prefix}.weight_hh_l0"], sd[f"{prefix}.bias_ih_l0"],
                      sd[f"{prefix}.bias_hh_l0"])
    b = gru_direction(x, sd[f"{prefix}.weight_ih_l0_reverse"], sd[f"{prefix}.weight_hh_l0_reverse"],
                      sd[f"{prefix}.bias_ih_l0_reverse"], sd[f"{prefix}.bias_hh_l0_reverse"], reverse=True)
    return torch.cat([f, b], dim=2)


def bigru_forward(sd, mels, ar=None, use_tanh=False, eps=1e-5):
    """BiGRU.forward in eval mode (models/pytorch_models.py:47-77).  mels (N, C, T) -> (N, C_out, T)."""
    x = mels
    if ar is not None:                                                        # :57-60
        a = past_fc_encoder(sd, ar)
        x = torch.cat((x, a.unsqueeze(2).repeat(1, 1, x.shape[2])), dim=1)
    h = x.transpose(1, 2)                                                     # (N, T, C)
    h = bigru_layer(sd, "gru1", h)
    h = bigru_layer(sd, "gru2", h)
    h = F.linear(h, sd["fc1.0.weight"], sd["fc1.0.bias"])                    # (N, T, 128)
    h = (h - sd["bn.running_mean"]) / torch.sqrt(sd["bn.running_var"] + eps) * sd["bn.weight"] + sd["bn.bias"]
    w2 = sd["fc2.0.weight"] if use_tanh else sd["fc2.weight"]
    b2 = sd["fc2.0.bias"] if use_tanh else sd["fc2.bias"]
    h = F.linear(h, w2, b2)
    if use_tanh:
        h = torch.tanh(h)
    return h.transpose(1, 2)


def bigru_inference(sd, c, normalize_before=True, mean=None, scale=None, **kw):
    """BiGRU.inference (:90-109): c (T, C) -> (T, C_out); statistics from register_stats when normalising."""
    if normalize_before:
        c = (c - mean) / scale
    return bigru_forward(sd, c.unsqueeze(0).transpose(1, 2), **kw).transpose(1, 2).squeeze(0)
