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
        # every parameter starts on a 128-byte boundary of the flat buffers: the weight-side kernels (prep / unprep /
        # weight-norm / their 128-bit row accesses) need 16-byte aligned rows, and a dense concatenation loses that after
        # the first 1-element bias (86 % of the discriminator's elements sat at odd offsets).  The gaps stay zero.
        self.offsets, total = {}, 0
        for n, p in named:
            self.offsets[n] = total
            total += (p.numel() + 31) // 32 * 32
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.wire = None
        self.views, self.grad_views = {}, {}
        with torch.no_grad():
            for n, p in named:
                k, off = p.numel(), self.offsets[n]
                view = self.flat[off:off + k].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = self.grad[off:off + k].view(p.shape)
                p.grad = gview
                self.views[n], self.grad_views[n] = view, gview
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
        """One Adam update from ``self.grad`` — or from ``self.wire`` when the data-parallel exchange left the reduced
        gradient in a (bf16) wire buffer (set by TrainStep)."""
        g = self.wire if self.wire is not None else self.grad
        call("artic_adam_step_wire", ptr(self.flat), ptr(g), _lib.DTYPE_CODE[g.dtype], ptr(self.m), ptr(self.v),
             self.flat.numel(), ptr(self.hyper))
        call("artic_adam_tick", ptr(self.hyper))
        if hasattr(self.module, "mark_weights_dirty"):
            self.module.mark_weights_dirty()

    # ---- checkpointing (bin/train.py:140-239 saves optimizer + scheduler state) -------
    def step_count(self):
        raw = self.hyper.cpu().numpy().tobytes()
        return AdamHyper.from_buffer_copy(raw).step

    def _current_lr(self, step):
        h = self._host_hyper
        n = sum(1 for i in range(h.n_milestones) if step >= h.milestones[i])
        return h.lr0 * (h.gamma ** n)

    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout (the reference saves ``optimizer.state_dict()``,
        bin/train.py:147-176): per-parameter ``state[i] = {step, exp_avg, exp_avg_sq}`` keyed by the index of the
        parameter in ``module.parameters()`` order, and one ``param_groups`` entry — so a checkpoint written here
        resumes in the reference and vice versa."""
        step = self.step_count()
        h = self._host_hyper
        m, v = self.m.cpu(), self.v.cpu()
        state = {}
        for i, (n, view) in enumerate(self.views.items()):
            k, off = view.numel(), self.offsets[n]
            if step > 0:
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": m[off:off + k].view(view.shape).clone(),
                            "exp_avg_sq": v[off:off + k].view(view.shape).clone()}
        group = {"lr": self._current_lr(step), "betas": (h.beta1, h.beta2), "eps": h.eps, "weight_decay": 0.0,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "initial_lr": h.lr0, "params": list(range(len(self.views)))}
        return {"state": state, "param_groups": [group]}

    def scheduler_state_dict(self):
        """``torch.optim.lr_scheduler.MultiStepLR.state_dict()`` layout (stepped once per iteration)."""
        from collections import Counter
        h = self._host_hyper
        step = self.step_count()
        return {"milestones": Counter(int(h.milestones[i]) for i in range(h.n_milestones)), "gamma": h.gamma,
                "base_lrs": [h.lr0], "last_epoch": step, "_step_count": step + 1, "verbose": False,
                "_get_lr_called_within_step": False, "_last_lr": [self._current_lr(step)]}

    def load_state_dict(self, sd, scheduler_sd=None):
        """Accepts ``torch.optim.Adam.state_dict()`` (reference checkpoints and the ones written here) or the flat
        layout of earlier versions of this package; raises on anything else instead of silently restarting the
        moments.  ``scheduler_sd`` (MultiStepLR state) supplies the step counter when the optimizer state holds
        none (no step taken yet)."""
        if "exp_avg" in sd and "step" in sd:                       # flat layout (round-1 checkpoints: dense concatenation)
            dense = sum(v.numel() for v in self.views.values())
            if sd["exp_avg"].numel() != dense:
                raise ValueError(f"flat optimizer state has {sd['exp_avg'].numel()} elements, this module has {dense}")
            m, v = torch.zeros_like(self.m, device="cpu"), torch.zeros_like(self.v, device="cpu")
            src = 0
            for n, view in self.views.items():
                k, off = view.numel(), self.offsets[n]
                m[off:off + k] = sd["exp_avg"].reshape(-1)[src:src + k].float().cpu()
                v[off:off + k] = sd["exp_avg_sq"].reshape(-1)[src:src + k].float().cpu()
                src += k
            self.m.copy_(m)
            self.v.copy_(v)
            step = int(sd["step"])
        elif "state" in sd and "param_groups" in sd:
            names = list(self.views.keys())
            n_params = len(sd["param_groups"][0]["params"]) if len(sd["param_groups"]) == 1 else -1
            if n_params != len(names):
                raise ValueError(f"optimizer state has {n_params} parameters in {len(sd['param_groups'])} group(s), "
                                 f"this module has {len(names)}")
            step = 0
            m, v = torch.zeros_like(self.m, device="cpu"), torch.zeros_like(self.v, device="cpu")
            for i, n in enumerate(names):
                view = self.views[n]
                k, off = view.numel(), self.offsets[n]
                st = sd["state"].get(i)
                if st is not None:
                    if tuple(st["exp_avg"].shape) != tuple(view.shape):
                        raise ValueError(f"optimizer state {i} has shape {tuple(st['exp_avg'].shape)}, parameter "
                                         f"{n} has {tuple(view.shape)} (parameter order differs)")
                    m[off:off + k] = st["exp_avg"].reshape(-1).float()
                    v[off:off + k] = st["exp_avg_sq"].reshape(-1).float()
                    step = max(step, int(float(st["step"])))
            self.m.copy_(m)
            self.v.copy_(v)
            g = sd["param_groups"][0]
            self._host_hyper.lr0 = float(g.get("initial_lr", self._host_hyper.lr0))
            self._host_hyper.beta1, self._host_hyper.beta2 = (float(b) for b in g["betas"])
            self._host_hyper.eps = float(g["eps"])
        else:
            raise ValueError("unknown optimizer state layout (expected torch.optim.Adam.state_dict())")
        if scheduler_sd is not None and "last_epoch" in scheduler_sd:
            step = max(step, int(scheduler_sd["last_epoch"]))
            if "gamma" in scheduler_sd:
                self._host_hyper.gamma = float(scheduler_sd["gamma"])
            ms = sorted(scheduler_sd.get("milestones", {}))
            if ms:
                assert len(ms) <= 8
                self._host_hyper.n_milestones = len(ms)
                for i, mval in enumerate(ms):
                    self._host_hyper.milestones[i] = int(mval)
        self._host_hyper.step = step
        self._upload()
