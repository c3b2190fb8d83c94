"""Fused Adam + MultiStepLR on flat fp32 buffers (reference: torch.optim.Adam with
MultiStepLR stepped every iteration, bin/train.py:372-383,424-435,1750-1789; yaml
egs/ema/voc1/conf/e2w_hifigan.yaml:142-169).

All parameters of a module are re-homed into ONE contiguous buffer (``param.data`` become
views), gradients into a second one — so the optimiser is a single kernel launch and the
data-parallel gradient exchange a single contiguous NCCL message.  Learning-rate schedule
and step counter live in device memory, which keeps a captured CUDA graph valid across steps.
"""
import ctypes

import torch

from . import _lib
from ._lib import AdamHyper, call, ptr


class FusedAdam:
    def __init__(self, module, lr=1e-4, betas=(0.5, 0.9), eps=1e-8, weight_decay=0.0,
                 gamma=0.5, milestones=(80000, 160000, 240000, 320000)):
        if weight_decay != 0.0:
            raise NotImplementedError("weight_decay != 0 is not on the hot path (yaml: 0.0)")
        assert len(milestones) <= 8
        self.module = module
        named = [(n, p) for n, p in module.named_parameters()]
        assert named, "module has no parameters"
        dev = named[0][1].device
        _lib.require_cuda(named[0][1], "parameters")
        total = sum(p.numel() for _, p in named)
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, self.grad_views = {}, {}
        off = 0
        with torch.no_grad():
            for n, p in named:
                k = p.numel()
                view = self.flat[off:off + k].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = self.grad[off:off + k].view(p.shape)
                p.grad = gview
                self.views[n], self.grad_views[n] = view, gview
                off += k
        h = AdamHyper()
        h.lr0, h.beta1, h.beta2, h.eps, h.gamma = lr, betas[0], betas[1], eps, gamma
        h.step, h.n_milestones = 0, len(milestones)
        for i, m in enumerate(milestones):
            h.milestones[i] = int(m)
        self._host_hyper = h
        self.hyper = torch.empty(ctypes.sizeof(AdamHyper), dtype=torch.uint8, device=dev)
        self._upload()
        if hasattr(module, "mark_weights_dirty"):
            module.mark_weights_dirty()

    def _upload(self):
        raw = bytes(self._host_hyper)
        self.hyper.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))

    def zero_grad(self):
        self.grad.zero_()

    def step(self):
        call("artic_adam_step", ptr(self.flat), ptr(self.grad), ptr(self.m), ptr(self.v), self.flat.numel(),
             ptr(self.hyper))
        call("artic_adam_tick", ptr(self.hyper))
        if hasattr(self.module, "mark_weights_dirty"):
            self.module.mark_weights_dirty()

    # ---- checkpointing (bin/train.py:140-239 saves optimizer + scheduler state) -------
    def step_count(self):
        raw = self.hyper.cpu().numpy().tobytes()
        return AdamHyper.from_buffer_copy(raw).step

    def state_dict(self):
        return {"step": self.step_count(), "exp_avg": self.m.cpu(), "exp_avg_sq": self.v.cpu(),
                "lr0": self._host_hyper.lr0, "betas": (self._host_hyper.beta1, self._host_hyper.beta2),
                "eps": self._host_hyper.eps, "gamma": self._host_hyper.gamma,
                "milestones": [self._host_hyper.milestones[i] for i in range(self._host_hyper.n_milestones)]}

    def load_state_dict(self, sd):
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        self._host_hyper.step = int(sd["step"])
        self._upload()
