"""Speech-to-EMA inversion encoder — the ``BiGRU`` plugin class (reference articulatory/models/pytorch_models.py:22-123),
inference forward on the B200 path (BASELINE configs[4]).

Same constructor keywords, ``state_dict`` keys (``gru1.weight_ih_l0`` ... ``fc1.0.weight``, ``bn.*``, ``fc2.weight`` /
``fc2.0.weight``, ``ar_model.model.{0,2,4,6,8}.*``), ``forward`` / ``inference`` / ``register_stats`` as the reference
class; the torch submodules are PARAMETER CONTAINERS only.  Schedule of one forward (eval mode: dropout = identity,
BatchNorm1d with running statistics):

  input assembly (N, C, T) -> channels-last (N, T, C [+ AR features])            artic_gen_input
  gi1 = x W_ih^T + b_ih for BOTH directions, all steps (one GEMM, tensor cores)  artic_tapconv (k = 1, bf16x3)
  gru1 recurrence, both directions                                               artic_bigru_layer (persistent cluster kernel)
  gi2, gru2 likewise
  fc1 -> BatchNorm1d(eval) -> fc2 folded into ONE affine map 2H -> out (exact algebra: no nonlinearity between them)
  [tanh]                                                                         torch.tanh on the (N, T, out) result

Training of this model (the reference trains it with bin/train.py and an L1 / mel loss) is outside the hot path:
``forward`` refuses to run with gradients enabled on its parameters.
"""
import logging

import numpy as np
import torch

from .. import _lib
from .._lib import F32, call, ptr
from ..convspec import ConvSpec
from ..engine import ConvLayer, SeqT, mlp_forward
from .hifigan import _PastFC


class BiGRU(torch.nn.Module):
    def __init__(self, in_channels=80, hidden_size=256, dropout=0.3, out_channels=1, use_ar=False, ar_input=512,
                 ar_hidden=256, ar_output=128, ar_channels=None, use_tanh=False, use_spk_emb=False, spk_emb_size=32,
                 spk_emb_hidden=32, precision="bf16x3"):
        super().__init__()
        if use_spk_emb:
            raise NotImplementedError("speaker-embedding conditioning is outside the B200 hot path")
        if hidden_size > 256 or hidden_size % 4:
            raise NotImplementedError("artic_bigru_layer keeps W_hh in the shared memory of a 4-CTA cluster: hidden_size must "
                                      "be a multiple of 4, <= 256")
        if precision not in ("bf16x3", "fp32"):
            raise ValueError("the inversion encoder runs in the fp32-accurate modes only (bf16x3 / fp32)")
        self.precision = precision
        self.hidden_size, self.in_channels, self.out_channels = hidden_size, in_channels, out_channels
        self.gru1 = torch.nn.GRU(input_size=in_channels, hidden_size=hidden_size, num_layers=1, batch_first=True,
                                 bidirectional=True)
        self.dropout1 = torch.nn.Dropout(dropout)
        self.gru2 = torch.nn.GRU(input_size=hidden_size * 2, hidden_size=hidden_size, num_layers=1, batch_first=True,
                                 bidirectional=True)
        self.dropout2 = torch.nn.Dropout(dropout)
        self.fc1 = torch.nn.Sequential(torch.nn.Linear(hidden_size * 2, 128), torch.nn.Dropout(p=dropout))
        self.bn = torch.nn.BatchNorm1d(128)
        self.use_tanh = use_tanh
        if not use_tanh:
            self.fc2 = torch.nn.Linear(128, out_channels)
        else:
            self.fc2 = torch.nn.Sequential(torch.nn.Linear(128, out_channels), torch.nn.Tanh())
        self.use_ar = use_ar
        self.ar_output = ar_output if use_ar else 0
        if use_ar:
            self.ar_model = _PastFC(ar_input, ar_hidden, ar_output)
        self._prep = None

    # ---- weight preparation (once per parameter version) ------------------------------------------------------
    def _key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _prepare(self):
        key = self._key()
        if self._prep is not None and self._prep["key"] == key:
            return self._prep
        dev = self.gru1.weight_ih_l0.device
        x3 = self.precision == "bf16x3"
        P = {"key": key}
        for name, gru in (("gru1", self.gru1), ("gru2", self.gru2)):
            cin = gru.input_size
            # both directions' input projections as ONE linear layer: rows [forward r|z|n, reverse r|z|n]
            w = torch.cat([gru.weight_ih_l0.data, gru.weight_ih_l0_reverse.data], 0).contiguous()
            b = torch.cat([gru.bias_ih_l0.data, gru.bias_ih_l0_reverse.data], 0).contiguous()
            lay = ConvLayer(ConvSpec("linear", cin, 6 * self.hidden_size), name, F32, F32, pad_in=True, x3=x3)
            lay.bind({f"{name}.weight": w, f"{name}.bias": b})
            lay.prep()
            P[name] = dict(lin=lay, keep=(w, b),
                           w_hh=torch.stack([gru.weight_hh_l0.data, gru.weight_hh_l0_reverse.data]).contiguous(),
                           b_hh=torch.stack([gru.bias_hh_l0.data, gru.bias_hh_l0_reverse.data]).contiguous())
        # fc1 -> BatchNorm1d (eval) -> fc2, folded in float64:  y = W2 (s * (W1 h + b1 - mu) + beta) + b2
        fc2 = self.fc2[0] if self.use_tanh else self.fc2
        s = (self.bn.weight.data.double() / torch.sqrt(self.bn.running_var.double() + self.bn.eps))
        w1, b1 = self.fc1[0].weight.data.double(), self.fc1[0].bias.data.double()
        w2, b2 = fc2.weight.data.double(), fc2.bias.data.double()
        wt = (w2 * s[None, :]) @ w1
        bt = w2 @ (s * (b1 - self.bn.running_mean.double()) + self.bn.bias.data.double()) + b2
        head = ConvLayer(ConvSpec("linear", 2 * self.hidden_size, self.out_channels), "head", F32, F32)
        hw, hb = wt.float().contiguous().to(dev), bt.float().contiguous().to(dev)
        head.bind({"head.weight": hw, "head.bias": hb})
        head.prep()
        P["head"] = dict(lin=head, keep=(hw, hb))
        if self.use_ar:
            lays = []
            for li in range(5):
                lin = self.ar_model.model[2 * li]
                lay = ConvLayer(ConvSpec("linear", lin.in_features, lin.out_features), f"ar{li}", F32, F32, x3=x3)
                lay.bind({f"ar{li}.weight": lin.weight.data, f"ar{li}.bias": lin.bias.data})
                lay.prep()
                lays.append(lay)
            P["ar"] = lays
        self._prep = P
        return P

    # ---- forward -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, mels, mask=None, spk_id=None, spk=None, ar=None, ph=None):
        """mels (N, C_mel, T) -> (N, C_out, T) (reference :47-77, eval mode)."""
        _lib.require_cuda(mels, "mels")
        if self.training:
            raise NotImplementedError("BiGRU on the B200 path is inference-only: call .eval() (training the inversion "
                                      "model is outside the hot path)")
        P = self._prepare()
        N, C, T = mels.shape
        dev, H = mels.device, self.hidden_size
        ar_feats, Ca = None, 0
        if self.use_ar:                                                          # :57-60
            Ca = self.ar_output
            acts = [SeqT.empty(N, 1, ar.numel() // N, F32, dev)] + [SeqT.empty(N, 1, lay.spec.cout, F32, dev) for lay in P["ar"]]
            mlp_forward(ar.reshape(N, -1).contiguous().float(), P["ar"], acts, F32, 0.1)
            h = acts[-1]
            ar_feats = h
        assert C + Ca == self.in_channels, f"in_channels {self.in_channels} != {C} + {Ca}"
        lin1 = P["gru1"]["lin"]
        x = SeqT.empty(N, T, lin1.kcig, F32, dev)                                 # zero-padded to the layer's width
        mc = mels.contiguous().float()
        call("artic_gen_input", ptr(mc), ptr(ar_feats.t) if ar_feats is not None else None, ptr(x.t), N, C, Ca,
             lin1.kcig, T, F32)
        for name in ("gru1", "gru2"):
            g = P[name]
            gi = SeqT.empty(N, T, 6 * H, F32, dev)
            g["lin"].forward(x, Y=gi, sp_y2=False)
            x = SeqT.empty(N, T, 2 * H, F32, dev)
            call("artic_bigru_layer", ptr(gi.t), ptr(g["w_hh"]), ptr(g["b_hh"]), ptr(x.t), N, T, H)
        y = SeqT.empty(N, T, self.out_channels, F32, dev)
        P["head"]["lin"].forward(x, Y=y)
        out = y.t.permute(0, 2, 1)
        return torch.tanh(out) if self.use_tanh else out

    def remove_weight_norm(self):
        """No layer of this model carries weight norm (the reference's method is a no-op sweep, :79-88)."""

    def inference(self, c, normalize_before=True, ar=None, spk=None):
        """c (T, in_channels) -> (T, out_channels) (reference :90-109)."""
        if len(c.shape) == 3:
            c = c.transpose(1, 2)
            c = c[0]
        if not isinstance(c, torch.Tensor):
            c = torch.tensor(c, dtype=torch.float).to(next(self.parameters()).device)
        if normalize_before:
            c = (c - self.mean) / self.scale
        c = self.forward(c.unsqueeze(0).transpose(1, 2), ar=ar, spk=spk)
        return c.transpose(1, 2).squeeze(0)

    def register_stats(self, stats):
        """Register (mean, scale) for input normalisation (reference :111-123)."""
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
