"""The fused G + D + spectral-loss train step (reference Trainer._train_step,
articulatory/bin/train.py:241-440), scheduled explicitly over the engines:

  G phase   y_ = G(x, ar); aux = [MR-STFT] + [mel]; p_ = D(cat(ar, y_)); p = D(cat(ar, y))
            gen_loss = l_aux*aux + l_adv*(adv(p_) + l_fm*fm(p_, p)); G <- Adam
            (D's weight gradients are NOT computed here: the reference computes and then
             discards them, bin/train.py:373 then :424)
  D phase   y_ = G_new(x, ar) (no grad); D(real) is REUSED from the G phase (D is not updated
            in between, so it is the same tensor the reference recomputes at :415);
            dis_loss = real + fake; D <- Adam

Differences from the reference that do not change results: no per-step ``.item()`` host
syncs (the nine logged scalars are accumulated on the device and read at the log
interval), one D(real) forward instead of two, no D wgrad in the G phase.  The steady-state
step is captured once in a CUDA graph (per rank) and replayed.
"""
import os as _os
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import F32, call, ptr
from .engine import SeqT, critical_stream, fork_join, slice_seq
from .losses.spectral import MelSpectrogramLoss, MultiResolutionSTFTLoss
from .optim import FusedAdam

#: ARTIC_BG="prep,spectral,order": which branches run on low-priority streams / in which order (see
#: engine.fork_join).  Measured on B200, 20-step runs, repeatable to +-0.02 ms (gpurun_out/r1_bg_66.log):
#: none 12.87, order 12.81 / 12.85, order+spectral 12.75 ms — the discriminator chain is enqueued before the
#: spectral losses and those run at low priority, so the ~1600 frame-FFT blocks no longer sit in front of the
#: chain's first (tiny) kernels.  "order2" (12.70): the generator forward is enqueued ahead of the big-grid weight
#: prep / gradient clears it overlaps with.  "prep" (weight re-materialisation in the background) does not pay.
_BG = set(filter(None, _os.environ.get("ARTIC_BG", "order,spectral,order2").split(",")))
#: ARTIC_SPLIT_DBWD=1: the discriminator's backward over the REAL half of [fake | real] does not depend on the generator
#: update, so it runs as a low-priority background branch under the generator backward of the G phase; the D phase then
#: only back-propagates the fake half before the (single) weight-gradient relayout.  Same results (parity tests run both
#: ways).  OFF by default: taking 1.7 ms of work off the critical path changed the step by 0.00 ms (12.430 vs 12.431 ms,
#: bf16; 28.27 vs 28.16 bf16x3) — the step is bound by the total SM time of its ~500 tensor-core launches, not by its
#: dependency chains.
_SPLIT_DBWD = _os.environ.get("ARTIC_SPLIT_DBWD", "0") == "1"

LOG_KEYS = ["train/spectral_convergence_loss", "train/log_stft_magnitude_loss", "train/mel_loss",
            "train/adversarial_loss", "train/feature_matching_loss", "train/generator_loss",
            "train/real_loss", "train/fake_loss", "train/discriminator_loss"]
_MEL, _ADV, _FM, _REAL, _FAKE = range(5)


def _event_on(stream):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


class TrainStep:
    def __init__(self, generator, discriminator, config: Dict, device, world_size=1, all_reduce=None, grad_wire=None):
        """``config`` uses the reference YAML keys (use_stft_loss, stft_loss_params, use_mel_loss,
        mel_loss_params, lambda_aux, lambda_adv, lambda_feat_match, *_optimizer_params,
        *_scheduler_params, *_train_start_steps).  ``all_reduce(flat_grad)`` (optional) sums a
        flat gradient buffer over data-parallel ranks; ``grad_wire`` (optional): see below."""
        self.G, self.D, self.cfg, self.dev = generator, discriminator, config, torch.device(device)
        self.world, self.all_reduce = world_size, all_reduce
        if config.get("generator_optimizer_type", "Adam") != "Adam" or \
                config.get("discriminator_optimizer_type", "Adam") != "Adam":
            raise NotImplementedError("the fused train step implements Adam (yaml default)")
        for k in ("generator", "discriminator"):
            if config.get(f"{k}_scheduler_type", "MultiStepLR") != "MultiStepLR":
                raise NotImplementedError("the fused train step implements MultiStepLR (yaml default)")
            if config.get(f"{k}_grad_norm", -1) > 0:
                raise NotImplementedError("gradient clipping is off in the shipped yamls")
        if not config.get("use_feat_match_loss", True):
            raise NotImplementedError("use_feat_match_loss=false is not on the hot path")
        self.use_stft = bool(config.get("use_stft_loss", False))
        self.use_mel = bool(config.get("use_mel_loss", False))
        self.stft = MultiResolutionSTFTLoss(**config.get("stft_loss_params", {})) if self.use_stft else None
        self.mel = MelSpectrogramLoss(**config["mel_loss_params"]).to(self.dev) if self.use_mel else None
        self.l_aux = float(config.get("lambda_aux", 1.0))
        self.l_adv = float(config.get("lambda_adv", 1.0))
        self.l_fm = float(config.get("lambda_feat_match", 1.0))
        fm = config.get("feat_match_loss_params", {})
        if fm.get("average_by_layers", True) or fm.get("average_by_discriminators", True) or \
                fm.get("include_final_outputs", False) or \
                config.get("generator_adv_loss_params", {}).get("average_by_discriminators", True) or \
                config.get("discriminator_adv_loss_params", {}).get("average_by_discriminators", True):
            raise NotImplementedError("only the yaml's un-averaged adversarial / feature-match sums are fused")
        self.g_start = config.get("generator_train_start_steps", 0)
        self.d_start = config.get("discriminator_train_start_steps", 0)

        def opt(module, key):
            op = dict(config.get(f"{key}_optimizer_params", {}))
            sp = config.get(f"{key}_scheduler_params", {})
            return FusedAdam(module, lr=op.get("lr", 1e-3), betas=tuple(op.get("betas", (0.9, 0.999))),
                             eps=op.get("eps", 1e-8), weight_decay=op.get("weight_decay", 0.0),
                             gamma=sp.get("gamma", 0.1), milestones=tuple(sp.get("milestones", ())))

        self.optG, self.optD = opt(generator, "generator"), opt(discriminator, "discriminator")
        if all_reduce is not None and grad_wire is not None:
            # ``grad_wire(flat)`` (DataParallel.wire_of): the buffer all_reduce leaves the reduced gradient in; Adam reads
            # it directly (bf16 wire: no cast back to the fp32 gradient buffer)
            self.optG.wire, self.optD.wire = grad_wire(self.optG.grad), grad_wire(self.optD.grad)
        R = len(self.stft.resolutions) if self.use_stft else 0
        self.R = R
        self.slots = torch.zeros(8, dtype=torch.float32, device=self.dev)
        self.stft_sums = torch.zeros((max(R, 1), 3), dtype=torch.float32, device=self.dev)
        self.stft_numel = torch.zeros(max(R, 1), dtype=torch.float32, device=self.dev)
        self.vals = torch.zeros(9, dtype=torch.float32, device=self.dev)
        self.running = torch.zeros(9, dtype=torch.float32, device=self.dev)
        self.steps = 0
        self._graph = None
        self._static = None
        self._overlap, self._ev_d, self._gfwd = False, None, None
        self._trace = None
        self.ar_len = generator._cfg["ar_input"] if generator.use_ar else 0

    # ------------------------------------------------------------------------------
    def _disc_input(self, ar, ys):
        """cat([ar, y], dim=2) for every y in ``ys`` stacked along the batch (bin/train.py:345-346):
        (len(ys) * B, 1, La + T) fp32."""
        B, _, T = ys[0].shape
        La = self.ar_len
        out = torch.empty((len(ys) * B, 1, La + T), dtype=torch.float32, device=self.dev)
        for i, y in enumerate(ys):
            call("artic_concat_time", ptr(ar) if La else None, ptr(y), ptr(out[i * B:]), B, La, T, La + T, F32)
        return out

    def _adv_seed(self, outs, target, slot, scale_w, grads=None, lo=0):
        """sum over discriminators of mean((logits - target)^2) -> slots[slot]; writes the logits
        gradients (scaled by scale_w) into batch items [lo, lo + B) of ``grads`` when given."""
        for ci, lst in enumerate(outs):
            lg = lst[-1]
            n = lg.numel()
            call("artic_sqerr_sum", ptr(lg.t), n, float(target), 1.0 / n, ptr(self.slots[slot:]), lg.code)
            if grads is not None:
                g = slice_seq(grads[ci], lo, lo + lg.N // lg.n_inner)
                call("artic_sqerr_bwd", ptr(lg.t), n, float(target), scale_w / n, ptr(g.t), 0, lg.code)

    # Both discriminator inputs of a phase travel as ONE batch [fake | real] (2B items): one set
    # of launches per layer, and in the D phase one backward / one weight gradient over both.
    def _phase_g_fwd(self, x, ar):
        """Generator forward alone (its own graph segment in the data-parallel schedule: it needs nothing from the
        discriminator, so it runs while the previous step's D gradients are still being exchanged)."""
        engG = self.G._ensure_ready()
        self._gfwd = (engG,) + tuple(engG.forward(x, ar, save=True))

    def _phase_g(self, x, y, ar, train_d_active, forwarded=False):
        G, D = self.G, self.D
        B, _, T = y.shape
        inv_w = 1.0 / self.world
        if forwarded:
            engG, y_, tapeG = self._gfwd
            self._gfwd = None
            engD = D._ensure_ready() if train_d_active else None     # no-op: re-materialised right after Adam(D)
        else:
            engG = G._ensure_ready()
            # the discriminator's weights (updated at the end of the previous step) are re-materialised on a
            # side stream while the generator forward runs
            if "order2" in _BG:      # generator forward enqueued first (side stream), the big-grid weight prep after it
                engD, (y_, tapeG) = fork_join([lambda: D._ensure_ready() if train_d_active else None,
                                               lambda: engG.forward(x, ar, save=True)])
            else:
                (y_, tapeG), engD = fork_join([lambda: engG.forward(x, ar, save=True),
                                               lambda: D._ensure_ready() if train_d_active else None],
                                              background=(1,) if "prep" in _BG else ())
        y2d, t2d = y_.reshape(B, T), y.reshape(B, T)
        dy = torch.zeros((B, 1, T), dtype=torch.float32, device=self.dev)
        self.slots.zero_()
        self.stft_sums.zero_()

        def spectral():
            # the STFT resolutions and the mel loss are independent of each other (dy is accumulated atomically)
            jobs = []
            if self.use_stft:                                                 # bin/train.py:289-297
                for r, res in enumerate(self.stft.resolutions):
                    def job(r=r, res=res):
                        res.forward(y2d, t2d, self.stft_sums[r])
                        res.backward(y2d, t2d, self.stft_sums[r], self.l_aux * inv_w / self.R,
                                     self.l_aux * inv_w / self.R, dy)
                    jobs.append(job)
            if self.use_mel:                                                  # :313-316
                def mel_job():
                    n = self.mel.numel(B, T)
                    self.mel.loss_and_grad(y2d, t2d, 1.0 / n, self.slots[_MEL:], self.l_aux * inv_w / n, dy)
                jobs.append(mel_job)
            if jobs:
                fork_join(jobs)

        tape2 = None
        if not train_d_active:
            spectral()
        else:                                                                 # :350-364
            return self._phase_g_adv(x, y, ar, y_, tapeG, engG, engD, dy, spectral)
        self.optG.zero_grad()
        engG.backward(tapeG, dy, self.optG.grad_views)
        return tape2

    def _phase_g_adv(self, x, y, ar, y_, tapeG, engG, engD, dy, spectral):
        """Adversarial + feature-matching part of the generator phase; runs concurrently with the
        spectral losses (both only read y_; they meet in dy)."""
        B, _, T = y.shape
        inv_w = 1.0 / self.world
        state = {}

        def adversarial():
            outs2, tape2 = engD.forward(None, save=True, parts=(ar if self.ar_len else None, (y_, y)))
            state["tape2"] = tape2
            outs_f = [[slice_seq(o, 0, B) for o in lst] for lst in outs2]
            outs_r = [[slice_seq(o, B, 2 * B) for o in lst] for lst in outs2]
            lg_grads = [lst[-1].like() for lst in outs_f]

            def seed_chain(ci):
                # adversarial seed of the logits + feature-matching term of every feature map of one
                # sub-discriminator (its own stream: ~6 small kernels per chain overlap across chains)
                self._adv_seed(outs_f[ci:ci + 1], 1.0, _ADV, self.l_adv * inv_w, lg_grads[ci:ci + 1])
                dl = []
                for a, b in zip(outs_f[ci][:-1], outs_r[ci][:-1]):
                    n = a.numel()
                    g = a.like()
                    if a.code == _lib.BF16:
                        call("artic_l1_sum_bwd", ptr(a.t), ptr(b.t), n, 1.0 / n, ptr(self.slots[_FM:]),
                             self.l_adv * self.l_fm * inv_w / n, ptr(g.t), a.code)
                    else:
                        call("artic_l1_sum", ptr(a.t), ptr(b.t), n, 1.0 / n, ptr(self.slots[_FM:]), a.code)
                        call("artic_l1_bwd", ptr(a.t), ptr(b.t), n, self.l_adv * self.l_fm * inv_w / n, ptr(g.t), 0, a.code)
                    dl.append(g)
                dl.append(lg_grads[ci])
                return dl

            douts = fork_join([lambda ci=ci: seed_chain(ci) for ci in range(len(outs_f))])
            state["d_in"] = engD.backward(engD.slice_tape(tape2, 0, B), douts, grads=None, need_dx=True)    # dgrad only

        # the discriminator chain is enqueued FIRST (side stream), the losses after it: the ~1600 frame-FFT
        # blocks of the mel loss otherwise fill every SM before the chain's first kernels get one
        if "order" in _BG:
            _, res = fork_join([spectral, adversarial], background=(0,) if "spectral" in _BG else ())
        else:
            fork_join([adversarial, spectral], background=(1,) if "spectral" in _BG else ())
        La = self.ar_len
        tape2 = state["tape2"]

        def g_backward():
            call("artic_add_rows", ptr(state["d_in"]) + 4 * La, La + T, ptr(dy), T, B, T)
            self.optG.zero_grad()
            engG.backward(tapeG, dy, self.optG.grad_views)

        def d_real_backward():
            # bin/train.py:415-418, real half: D(real) and D's weights are final here (D is updated at the very end of the
            # step), so its share of dL_D/dθ_D is computed now, in the background of the generator backward
            self.optD.zero_grad()
            engD.wset.zero()
            real = [[slice_seq(acts[-1], B, 2 * B)] for acts in tape2["chains"]]
            lg = [lst[0].like() for lst in real]
            fork_join([lambda ci=ci: self._adv_seed(real[ci:ci + 1], 1.0, _REAL, inv_w, lg[ci:ci + 1]) for ci in range(len(real))])
            douts = [[None] * (len(acts) - 2) + [lg[ci]] for ci, acts in enumerate(tape2["chains"])]
            engD.backward(engD.slice_tape(tape2, B, 2 * B), douts, grads=self.optD.grad_views, need_dx=False, pre_zeroed=True,
                          finalize=False)

        self._d_real_done = False
        if _SPLIT_DBWD:
            fork_join([g_backward, d_real_backward], background=(1,))
            self._d_real_done = True
        else:
            g_backward()
        return tape2

    def _phase_d(self, x, y, ar, tape2):
        G, D = self.G, self.D
        engG = G._ensure_ready()          # re-materialises the updated generator weights
        engD = D._ensure_ready()
        B = y.shape[0]
        inv_w = 1.0 / self.world
        def clear_d_grads():      # two 283 MB fills: under the generator forward instead of in front of the D backward
            self.optD.zero_grad()
            engD.wset.zero()

        real_done = tape2 is not None and getattr(self, "_d_real_done", False)
        self._d_real_done = False
        if real_done:              # the gradient buffers already hold the real half's share (G phase): do not clear them
            y_, _ = engG.forward(x, ar, save=False)                                            # bin/train.py:390-400
        elif "order2" in _BG:
            _, (y_, _) = fork_join([clear_d_grads, lambda: engG.forward(x, ar, save=False)])
        else:
            (y_, _), _ = fork_join([lambda: engG.forward(x, ar, save=False), clear_d_grads])   # bin/train.py:390-400
        if tape2 is None:
            _, tape2 = engD.forward(None, save=True, parts=(ar if self.ar_len else None, (y_, y)))
        else:
            # D(real) is REUSED from the G phase; D(fake) overwrites the stale fake half in place
            engD.forward(None, save=True, into=tape2, lo=0, parts=(ar if self.ar_len else None, (y_,)))
        if real_done:
            fake = [[slice_seq(acts[-1], 0, B)] for acts in tape2["chains"]]
            lg = [lst[0].like() for lst in fake]
            fork_join([lambda ci=ci: self._adv_seed(fake[ci:ci + 1], 0.0, _FAKE, inv_w, lg[ci:ci + 1]) for ci in range(len(fake))])
            douts = [[None] * (len(acts) - 2) + [lg[ci]] for ci, acts in enumerate(tape2["chains"])]
            engD.backward(engD.slice_tape(tape2, 0, B), douts, grads=self.optD.grad_views, need_dx=False, pre_zeroed=True)
            return
        outs2 = [acts[1:] for acts in tape2["chains"]]
        outs_f = [[slice_seq(lst[-1], 0, B)] for lst in outs2]
        outs_r = [[slice_seq(lst[-1], B, 2 * B)] for lst in outs2]
        lg_grads = [lst[-1].like() for lst in outs2]
        def seed_chain(ci):                                                   # :415-418, one stream per chain
            self._adv_seed(outs_r[ci:ci + 1], 1.0, _REAL, inv_w, lg_grads[ci:ci + 1], lo=B)
            self._adv_seed(outs_f[ci:ci + 1], 0.0, _FAKE, inv_w, lg_grads[ci:ci + 1], lo=0)

        fork_join([lambda ci=ci: seed_chain(ci) for ci in range(len(outs2))])
        douts = [[None] * (len(lst) - 1) + [lg_grads[ci]] for ci, lst in enumerate(outs2)]
        engD.backward(tape2, douts, grads=self.optD.grad_views, need_dx=False, pre_zeroed=True)

    # The step is cut into three segments so that the (optional) data-parallel gradient
    # all-reduce can run between CUDA-graph replays:  seg1 = G phase up to dL/dθ_G,
    # seg2 = Adam(G) + D phase up to dL/dθ_D,  seg3 = Adam(D) + log assembly.
    def _seg1(self, x, y, ar, steps, forwarded=False):
        self._real = None
        if steps > self.g_start:
            self._real = self._phase_g(x, y, ar, steps > self.d_start, forwarded=forwarded)
        else:
            self.slots.zero_()
            self.stft_sums.zero_()

    def _seg2(self, x, y, ar, steps):
        if steps > self.g_start:
            self.optG.step()
        if steps > self.d_start:
            self._phase_d(x, y, ar, self._real)
        self._real = None

    def _seg3(self, steps, prep_d=False):
        if steps > self.d_start:
            self.optD.step()
        call("artic_train_log", ptr(self.slots), ptr(self.stft_sums), ptr(self.stft_numel),
             self.R if steps > self.g_start else 0, self.l_aux, self.l_adv, self.l_fm, ptr(self.vals),
             ptr(self.running))
        if prep_d:          # overlapped schedule: D's weights are re-materialised here, off the critical path
            self.D._ensure_ready()

    def _step_impl(self, x, y, ar, steps):
        """One reference train step at global step ``steps`` (gates as bin/train.py:268,350,388)."""
        self._seg1(x, y, ar, steps)
        if self.all_reduce is not None and steps > self.g_start:
            self.all_reduce(self.optG.grad)
        self._seg2(x, y, ar, steps)
        if self.all_reduce is not None and steps > self.d_start:
            self.all_reduce(self.optD.grad)
        self._seg3(steps)

    # ------------------------------------------------------------------------------
    def step(self, x, y, ar, use_graph=True):
        """One train step.  x (B, C, T'), y (B, 1, T), ar (B, 1, ar_len) fp32; CUDA tensors, or
        (pinned) host tensors which are copied host->device inside the step."""
        steady = self.steps > max(self.g_start, self.d_start)
        self._ensure_numel(y.shape[0], y.shape[2])
        if ar is None:      # non-AR generator (use_ar: false): the collater emits no 'ar'; a 0-length context stands in
            assert self.ar_len == 0, "this generator is autoregressive: the batch must carry 'ar'"
            ar = y.new_zeros((y.shape[0], 1, 0))
        if use_graph and steady:
            if self._graph is None or self._static[0].shape != x.shape:
                self._capture(x, y, ar)
            for dst, src in zip(self._static, (x, y, ar)):
                dst.copy_(src, non_blocking=True)
            if self._overlap:
                # data-parallel schedule (four graphs): the exchange of the D gradients, Adam(D) and D's weight
                # re-materialisation of step k run on the exchange stream UNDER the generator forward of step k + 1;
                # only the (5x smaller) exchange of the G gradients is on the critical path
                g1a, g1b, g2, g3 = self._graph
                main = torch.cuda.current_stream()
                tr = self._trace                     # tools/dp_timeline.py: CUDA events at the schedule's joints
                mark = (lambda st: tr[-1].append(_event_on(st))) if tr is not None else (lambda st: None)
                if tr is not None:
                    tr.append([])
                mark(main)
                g1a.replay()
                mark(main)
                main.wait_event(self._ev_d)
                g1b.replay()
                mark(main)
                if self.all_reduce is not None:
                    self.all_reduce(self.optG.grad)
                mark(main)
                g2.replay()
                self._ev_2.record(main)
                mark(main)
                with torch.cuda.stream(self._xs):
                    self._xs.wait_event(self._ev_2)
                    if self.all_reduce is not None:
                        self.all_reduce(self.optD.grad)
                    mark(self._xs)
                    g3.replay()
                    self._ev_d.record(self._xs)
                    mark(self._xs)
            else:
                g1, g2, g3 = self._graph
                g1.replay()
                if self.all_reduce is not None:
                    self.all_reduce(self.optG.grad)
                g2.replay()
                if self.all_reduce is not None:
                    self.all_reduce(self.optD.grad)
                g3.replay()
            # the replays re-wrote the flat weights (Adam) behind the modules' host-side caches: a later EAGER use
            # (eval_step, checkpoint-time inference) must re-materialise the prepared weights
            self.G.mark_weights_dirty()
            self.D.mark_weights_dirty()
        else:
            self.sync_exchange()
            x, y, ar = (t.to(self.dev, non_blocking=True).float().contiguous() for t in (x, y, ar))
            self._step_impl(x, y, ar, self.steps)
        self.steps += 1

    def sync_exchange(self):
        """Make the current stream wait for the overlapped tail of the last step (D gradient exchange, Adam(D), log
        assembly); a no-op in the single-GPU / non-overlapped schedules.  Called by every reader of the logged values
        and of D's parameters."""
        if getattr(self, "_overlap", False) and self._ev_d is not None:
            torch.cuda.current_stream().wait_event(self._ev_d)

    def values_tensor(self):
        """The nine logged scalars of the last step as a device tensor that is safe to read on the current stream."""
        self.sync_exchange()
        return self.vals

    def _ensure_numel(self, B, T):
        """Element counts of the STFT magnitude tensors (log-magnitude mean), kept on the device."""
        if self.use_stft and getattr(self, "_numel_for", None) != (B, T):
            host = torch.tensor([float(r.numel(B, T)) for r in self.stft.resolutions], dtype=torch.float32)
            self.stft_numel.copy_(host.to(self.dev))
            self._numel_for = (B, T)

    def _capture(self, x, y, ar):
        sx, sy, sa = (t.to(self.dev).clone().float().contiguous() for t in (x, y, ar))
        torch.cuda.synchronize()
        snap = self._snapshot()
        # one eager steady-state step on a side stream (lazy allocations, smem attributes, ...)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        ar_fn, self.all_reduce = self.all_reduce, None
        with torch.cuda.stream(s):
            self._step_impl(sx, sy, sa, self.steps)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        pool = torch.cuda.graph_pool_handle()
        graphs = []
        cap = critical_stream(self.dev)         # high priority: see engine.fork_join
        # (ARTIC_TAIL_OVERLAP=1 runs the same four-graph schedule on one GPU: Adam(D) + D weight re-materialisation of
        # step k under the generator forward of step k + 1)
        self._overlap = (ar_fn is not None and _os.environ.get("ARTIC_DP_OVERLAP", "1") != "0") or \
            _os.environ.get("ARTIC_TAIL_OVERLAP", "0") == "1"
        # G's prepared weights are refreshed inside the step (right after Adam(G)); materialise them now so that the
        # generator phase does not capture a second, redundant re-materialisation per step
        self.G._ensure_ready()
        if self._overlap:
            # D's prepared weights are refreshed by the tail segment (after Adam(D)); make them current before the
            # first replay and keep the re-materialisation out of the generator-phase graph
            self.D._ensure_ready()
            segs = (lambda: self._phase_g_fwd(sx, sa), lambda: self._seg1(sx, sy, sa, self.steps, forwarded=True),
                    lambda: self._seg2(sx, sy, sa, self.steps), lambda: self._seg3(self.steps, prep_d=True))
        else:
            segs = (lambda: self._seg1(sx, sy, sa, self.steps), lambda: self._seg2(sx, sy, sa, self.steps),
                    lambda: self._seg3(self.steps))
        for seg in segs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=cap):
                seg()
            graphs.append(g)
        self.all_reduce = ar_fn
        self._restore(snap)
        self.G._ensure_ready()
        if self._overlap:
            self.D._ensure_ready()
            self._xs = torch.cuda.Stream(device=self.dev, priority=-1)
            self._ev_2, self._ev_d = torch.cuda.Event(), torch.cuda.Event()
            self._ev_d.record(torch.cuda.current_stream())
        self._graph, self._static = tuple(graphs), (sx, sy, sa)

    def _snapshot(self):
        return [t.clone() for t in (self.optG.flat, self.optG.m, self.optG.v, self.optG.hyper,
                                    self.optD.flat, self.optD.m, self.optD.v, self.optD.hyper, self.running)]

    def _restore(self, snap):
        for dst, src in zip((self.optG.flat, self.optG.m, self.optG.v, self.optG.hyper,
                             self.optD.flat, self.optD.m, self.optD.v, self.optD.hyper, self.running), snap):
            dst.copy_(src)
        self.G.mark_weights_dirty()
        self.D.mark_weights_dirty()

    # ------------------------------------------------------------------------------
    EVAL_KEYS = [k.replace("train/", "eval/") for k in LOG_KEYS]

    @torch.no_grad()
    def eval_step(self, x, y, ar):
        """One evaluation step (reference Trainer._eval_step, bin/train.py:470-603): the nine losses of the train
        step on a dev batch, no gradients, no updates.  Returns a dict with the reference's ``eval/*`` keys
        (one host read); ``eval_running`` accumulates them for ``_eval_epoch``-style averaging.  D(real) and
        D(fake) are evaluated once and feed both the generator-side and the discriminator-side terms (the
        reference evaluates each twice, :572-587, with identical results under no_grad)."""
        if ar is None:
            assert self.ar_len == 0, "this generator is autoregressive: the batch must carry 'ar'"
            ar = y.new_zeros((y.shape[0], 1, 0))
        self.sync_exchange()
        x, y, ar = (t.to(self.dev, non_blocking=True).float().contiguous() for t in (x, y, ar))
        self._ensure_numel(y.shape[0], y.shape[2])
        if not hasattr(self, "eval_vals"):
            self.eval_vals = torch.zeros(9, dtype=torch.float32, device=self.dev)
            self.eval_running = torch.zeros(9, dtype=torch.float32, device=self.dev)
        engG, engD = self.G._ensure_ready(), self.D._ensure_ready()
        B, _, T = y.shape
        y_, _ = engG.forward(x, ar, save=False)
        y2d, t2d = y_.reshape(B, T), y.reshape(B, T)
        self.slots.zero_()
        self.stft_sums.zero_()
        if self.use_stft:                                                     # :515-520
            for r, res in enumerate(self.stft.resolutions):
                res.forward(y2d, t2d, self.stft_sums[r])
        if self.use_mel:                                                      # :536-540
            self.mel.accumulate(y2d, t2d, 1.0 / self.mel.numel(B, T), self.slots[_MEL:])
        outs2, _ = engD.forward(None, save=False, parts=(ar if self.ar_len else None, (y_, y)))    # :572-587, [fake | real]
        outs_f = [[slice_seq(o, 0, B) for o in lst] for lst in outs2]
        outs_r = [[slice_seq(o, B, 2 * B) for o in lst] for lst in outs2]
        self._adv_seed(outs_f, 1.0, _ADV, 0.0)                                # gen_adv(p_)
        for lf, lr in zip(outs_f, outs_r):                                    # feat_match(p_, p)
            for a, b in zip(lf[:-1], lr[:-1]):
                call("artic_l1_sum", ptr(a.t), ptr(b.t), a.numel(), 1.0 / a.numel(), ptr(self.slots[_FM:]), a.code)
        self._adv_seed(outs_r, 1.0, _REAL, 0.0)                               # dis_adv(p_, p)
        self._adv_seed(outs_f, 0.0, _FAKE, 0.0)
        call("artic_train_log", ptr(self.slots), ptr(self.stft_sums), ptr(self.stft_numel), self.R, self.l_aux,
             self.l_adv, self.l_fm, ptr(self.eval_vals), ptr(self.eval_running))
        return dict(zip(self.EVAL_KEYS, self.eval_vals.cpu().tolist()))

    def read_eval_logs(self, n_steps, reset=True):
        """Averages of the accumulated eval losses over ``n_steps`` eval steps (bin/train.py:624-629)."""
        vals = (self.eval_running / max(n_steps, 1)).cpu().tolist()
        if reset:
            self.eval_running.zero_()
        return dict(zip(self.EVAL_KEYS, vals))

    def read_logs(self, reset=True):
        """Device -> host read of the running sums (the reference's total_train_loss)."""
        self.sync_exchange()
        vals = self.running.cpu().tolist()
        if reset:
            self.running.zero_()
        return dict(zip(LOG_KEYS, vals))

    def last_values(self):
        self.sync_exchange()
        return dict(zip(LOG_KEYS, self.vals.cpu().tolist()))
