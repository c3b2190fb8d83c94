"""Data-parallel plumbing of the train step: the batch is sharded by utterance, G / D / both
Adam states are replicated, and the ONLY exchange is the sum of the flat gradient buffers
(reference intent: DistributedSampler + DDP, bin/train.py:1610-1646,1790-1801 — disabled there).

``torch.distributed`` carries the collective: NCCL over NVLink / NVSwitch on the GPU box, gloo in
the CPU tests.  The 1/world factor is folded into the loss seeds by ``TrainStep(world_size=...)``,
so the all-reduce is a plain sum and the optimiser sees the average of the per-rank gradients —
exactly what DDP would have produced.  Spectral convergence is a per-rank batch ratio.
"""
import os
from typing import Dict

import torch
import torch.distributed as dist


_SKIP = os.environ.get("ARTIC_DP_SKIP", "")       # "all" | "big" | "small": what-if switch for scaling experiments


def env_world():
    """(rank, local_rank, world_size) from the launcher's environment (1 process = defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


class DataParallel:
    def __init__(self, backend=None, device=None, compress=None):
        """``compress="bf16"``: gradients cross the wire as bf16 (half the bytes; DDP's ``bf16_compress_hook``
        semantics: cast, sum in bf16, cast back) — meant for the bf16 speed mode, whose gradients carry bf16 noise
        anyway.  Default: exact fp32 sum."""
        self.rank, self.local_rank, self.world = env_world()
        self.device = device
        self.compress = compress
        self._wire, self._claimed = {}, set()
        self.bytes_reduced = 0
        if self.world > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if (device is not None and torch.device(device).type == "cuda") else "gloo")
            kw = {"device_id": torch.device(device)} if backend == "nccl" else {}
            dist.init_process_group(backend, init_method="env://", **kw)

    # ---- parameters -------------------------------------------------------------------
    def broadcast_parameters(self, *modules, src=0):
        """Rank ``src``'s weights (and buffers) become everyone's initial weights."""
        if self.world == 1:
            return
        for m in modules:
            for t in list(m.parameters()) + list(m.buffers()):
                dist.broadcast(t.data, src)
            if hasattr(m, "mark_weights_dirty"):
                m.mark_weights_dirty()

    # ---- gradients --------------------------------------------------------------------
    def wire_of(self, flat: torch.Tensor):
        """The bf16 wire buffer of a flat gradient buffer (None for the exact fp32 exchange).  A caller that asks for it
        CONSUMES the reduced gradient from it (``FusedAdam.wire``): ``all_reduce(flat)`` then skips the cast back to
        fp32 — one 6-byte-per-parameter pass less on the exchange path."""
        if self.world == 1 or self.compress != "bf16":
            return None
        w = self._wire.get(flat.data_ptr())
        if w is None:
            w = self._wire[flat.data_ptr()] = torch.empty_like(flat, dtype=torch.bfloat16)
        self._claimed.add(flat.data_ptr())
        return w

    def all_reduce(self, flat: torch.Tensor):
        """Sum a flat gradient buffer over ranks, in place (called between the step's graph segments)."""
        if self.world > 1:
            if _SKIP and (_SKIP == "all" or (_SKIP == "big") == (flat.numel() > 30_000_000)):
                return                      # timing experiments only (ARTIC_DP_SKIP): the step's results are wrong
            if self.compress == "bf16":
                w = self._wire.get(flat.data_ptr())
                if w is None:
                    w = self._wire[flat.data_ptr()] = torch.empty_like(flat, dtype=torch.bfloat16)
                w.copy_(flat)
                dist.all_reduce(w, op=dist.ReduceOp.SUM)
                if flat.data_ptr() not in self._claimed:
                    flat.copy_(w)
                self.bytes_reduced += w.numel() * 2
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                self.bytes_reduced += flat.numel() * flat.element_size()

    def describe(self):
        """What carried the gradient exchange (for the bench line): backend, wire dtype, and — when NCCL_DEBUG=INFO
        was routed to a file by the caller (NCCL_DEBUG_FILE) — whether NCCL set up NVLS (in-switch reduction)."""
        if self.world == 1:
            return None
        d = {"backend": dist.get_backend(), "world": self.world, "wire_dtype": self.compress or "fp32"}
        try:
            d["nccl_version"] = ".".join(map(str, torch.cuda.nccl.version()))
        except Exception:
            pass
        path = os.environ.get("NCCL_DEBUG_FILE", "")
        if path:
            path = path.replace("%h", os.uname().nodename).replace("%p", str(os.getpid()))
            try:
                txt = open(path, errors="ignore").read()
                d["nvls"] = "NVLS" in txt
                algos = sorted({a for a in ("NVLS", "NVLSTree", "Ring", "Tree", "CollNet") if f" {a} " in txt or f"{a}/" in txt})
                d["nccl_log_mentions"] = algos
            except OSError:
                pass
        return d

    # ---- data -------------------------------------------------------------------------
    def shard(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """This rank's utterances of a global batch (contiguous split, ragged tail to the low ranks)."""
        out = {}
        for k, v in batch.items():
            n = v.shape[0]
            base, extra = divmod(n, self.world)
            lo = self.rank * base + min(self.rank, extra)
            hi = lo + base + (1 if self.rank < extra else 0)
            out[k] = v[lo:hi]
        return out

    def sampler_indices(self, n_items: int, epoch: int, shuffle=True):
        """Indices of this rank for one epoch (DistributedSampler semantics: seeded by the epoch,
        padded to a multiple of the world size by wrapping around)."""
        g = torch.Generator().manual_seed(epoch)
        idx = torch.randperm(n_items, generator=g).tolist() if shuffle else list(range(n_items))
        total = -(-n_items // self.world) * self.world
        idx += idx[: total - n_items]
        return idx[self.rank:total:self.world]

    # ---- logging ----------------------------------------------------------------------
    def mean_scalars(self, values: torch.Tensor) -> torch.Tensor:
        """Average a small tensor of logged scalars over ranks (one collective per log interval)."""
        if self.world > 1:
            values = values.clone()
            dist.all_reduce(values, op=dist.ReduceOp.SUM)
            values /= self.world
        return values

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    def close(self):
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()
