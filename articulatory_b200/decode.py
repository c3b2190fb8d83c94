"""Chunked autoregressive (HiFi-CAR) decoding on the B200 path.

The reference decodes one utterance at a time (bin/decode.py:31-83): the feature sequence is cut
into chunks of ``batch_max_steps // hop_size`` frames, every chunk is one generator forward
conditioned on the last ``ar_input`` samples of the previous chunk's output.  Chunks of one
utterance are strictly sequential, utterances are independent — so a batch of utterances
advances in lock-step, one generator forward per chunk index for the whole batch.  A chunk is
~80 tiny convolutions, i.e. launch-latency bound: the per-chunk forward (including the update of
the AR context) is captured ONCE per (batch, chunk length) in a CUDA graph and replayed.

Frame / sample indexing is the reference's, bit-exactly (tests/golden/indexing.json).
"""
from typing import List, Sequence

import torch


def chunk_plan(n_frames: int, batch_max_steps: int, hop_size: int):
    """[(frame_lo, frame_hi, sample_lo, sample_hi)] of reference bin/decode.py:45-56."""
    chunk = int(batch_max_steps / hop_size)
    return [(i, min(i + chunk, n_frames), i * hop_size, min(i + chunk, n_frames) * hop_size)
            for i in range(0, n_frames, chunk)]


class BatchedARDecoder:
    """Lock-step chunked AR decoding of a batch of utterances with one CUDA graph per chunk shape.

    ``model``: a ``HiFiGANGenerator`` with ``use_ar=True`` on a CUDA device (weight norm may be
    removed or not).  ``config``: the reference YAML dict (``batch_max_steps``, ``hop_size``,
    ``generator_params``)."""

    def __init__(self, model, config, use_graph=True):
        gp = config["generator_params"]
        self.model, self.use_graph = model, use_graph
        self.batch_max_steps, self.hop = config["batch_max_steps"], config["hop_size"]
        self.out_channels = gp.get("out_channels", 1)
        if self.out_channels != 1:
            raise NotImplementedError("multi-band (out_channels > 1) decoding is outside the B200 hot path")
        self.past = gp["ar_input"]
        self.chunk_frames = int(self.batch_max_steps / self.hop)
        self.dev = next(model.parameters()).device
        self._graphs = {}

    # -- one chunk for the whole batch ------------------------------------------------
    def _step_eager(self, cin, prev):
        cout = self.model(cin, ar=prev)                                   # (B, 1, frames*hop)
        n = cout.shape[2]
        if self.past <= self.batch_max_steps:                             # decode.py:77-78
            if n >= self.past:
                new_prev = cout[:, :, -self.past:]
            else:   # short last chunk: the reference slices a shorter context that is never used
                new_prev = torch.cat([prev[:, :, n:], cout], dim=2)
        else:                                                             # decode.py:79-81 (shift register)
            new_prev = torch.cat([prev[:, :, n:], cout], dim=2)
        return cout, new_prev

    def _graph_for(self, B, C, frames):
        key = (B, C, frames)
        g = self._graphs.get(key)
        if g is None:
            s_cin = torch.zeros((B, C, frames), dtype=torch.float32, device=self.dev)
            s_prev = torch.zeros((B, 1, self.past), dtype=torch.float32, device=self.dev)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                                 # warm-up: lazy allocations, smem attributes
                self._step_eager(s_cin, s_prev)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                cout, new_prev = self._step_eager(s_cin, s_prev)
                s_prev.copy_(new_prev)
            g = (graph, s_cin, s_prev, cout)
            self._graphs[key] = g
        return g

    @torch.no_grad()
    def decode(self, feats: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """feats: list of (T'_i, C) tensors (host or device).  Returns a list of (T'_i * hop,) waveforms,
        each identical to the reference ``ar_loop`` run on that utterance alone."""
        feats = [f.to(self.dev, dtype=torch.float32) for f in feats]
        C = feats[0].shape[1]
        outs = [torch.empty(len(f) * self.hop, dtype=torch.float32, device=self.dev) for f in feats]
        max_frames = max(len(f) for f in feats)
        prev = torch.zeros((len(feats), 1, self.past), dtype=torch.float32, device=self.dev)   # decode.py:59
        F = self.chunk_frames
        for lo in range(0, max_frames, F):
            # group the utterances still running by the length of their chunk at this index
            groups = {}
            for i, f in enumerate(feats):
                n = min(len(f) - lo, F)
                if n > 0:
                    groups.setdefault(n, []).append(i)
            for n, idx in groups.items():
                cin = torch.stack([feats[i][lo:lo + n] for i in idx]).permute(0, 2, 1).contiguous()   # (b, C, n)
                sel = torch.tensor(idx, device=self.dev)
                pv = prev.index_select(0, sel)
                if self.use_graph:
                    graph, s_cin, s_prev, s_out = self._graph_for(len(idx), C, n)
                    s_cin.copy_(cin)
                    s_prev.copy_(pv)
                    graph.replay()
                    cout, new_prev = s_out, s_prev
                else:
                    cout, new_prev = self._step_eager(cin, pv)
                prev.index_copy_(0, sel, new_prev)
                for j, i in enumerate(idx):
                    outs[i][lo * self.hop:(lo + n) * self.hop] = cout[j, 0]
        return outs
