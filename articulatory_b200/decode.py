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

from .engine import weight_multicast


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
            with torch.cuda.stream(side), weight_multicast(2):            # warm-up: lazy allocations, smem attributes
                self._step_eager(s_cin, s_prev)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            # few rows per chunk and three chains in flight: pairs of CTAs share the streamed weight tiles (TMA multicast)
            with torch.cuda.graph(graph), weight_multicast(2):
                cout, new_prev = self._step_eager(s_cin, s_prev)
                s_prev.copy_(new_prev)
            g = (graph, s_cin, s_prev, cout)
            self._graphs[key] = g
        return g

    @torch.no_grad()
    def decode(self, feats: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """feats: list of (T'_i, C) tensors (host or device).  Returns a list of (T'_i * hop,) waveforms,
        each identical to the reference ``ar_loop`` run on that utterance alone.

        Host-side work per chunk is two strided copies around the graph replay: the features live in ONE padded
        (B, C, T'max) device tensor, the waveforms in ONE (B, T'max * hop) buffer, and while every utterance is still
        running (the common case: equal lengths) the AR context never leaves the graph's static buffer."""
        B = len(feats)
        C = feats[0].shape[1]
        lens = [len(f) for f in feats]
        max_frames = max(lens)
        # one H2D copy per utterance, straight into the padded channel-first layout
        fpad = torch.zeros((B, C, max_frames), dtype=torch.float32, device=self.dev)
        for i, f in enumerate(feats):
            fpad[i, :, :lens[i]].copy_(f.to(self.dev, dtype=torch.float32, non_blocking=True).t())
        obuf = torch.empty((B, max_frames * self.hop), dtype=torch.float32, device=self.dev)
        prev = torch.zeros((B, 1, self.past), dtype=torch.float32, device=self.dev)   # decode.py:59
        F = self.chunk_frames
        in_graph = None         # (graph tuple) whose static AR buffer currently holds `prev` for ALL utterances
        for lo in range(0, max_frames, F):
            # group the utterances still running by the length of their chunk at this index
            groups = {}
            for i, n_i in enumerate(lens):
                n = min(n_i - lo, F)
                if n > 0:
                    groups.setdefault(n, []).append(i)
            for n, idx in groups.items():
                whole = len(idx) == B
                if self.use_graph:
                    g = self._graph_for(len(idx), C, n)
                    graph, s_cin, s_prev, s_out = g
                    if whole:
                        s_cin.copy_(fpad[:, :, lo:lo + n])
                        if in_graph is not g:
                            s_prev.copy_(prev if in_graph is None else in_graph[2])
                        graph.replay()
                        in_graph = g
                        obuf[:, lo * self.hop:(lo + n) * self.hop].copy_(s_out[:, 0])
                        continue
                    if in_graph is not None:            # leave the all-utterances fast path: materialise the context
                        prev.copy_(in_graph[2])
                        in_graph = None
                    sel = torch.tensor(idx, device=self.dev)
                    s_cin.copy_(fpad.index_select(0, sel)[:, :, lo:lo + n])
                    s_prev.copy_(prev.index_select(0, sel))
                    graph.replay()
                    cout, new_prev = s_out, s_prev
                else:
                    sel = torch.tensor(idx, device=self.dev)
                    cout, new_prev = self._step_eager(fpad.index_select(0, sel)[:, :, lo:lo + n].contiguous(),
                                                      prev.index_select(0, sel))
                prev.index_copy_(0, sel, new_prev)
                obuf[:, lo * self.hop:(lo + n) * self.hop].index_copy_(0, sel, cout[:, 0])
        return [obuf[i, :lens[i] * self.hop] for i in range(B)]
