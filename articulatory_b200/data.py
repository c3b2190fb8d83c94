"""Host-side batch assembly on the hot path's input side.

``SpeechCollater`` mirrors reference bin/train.py:865-1098 for the configuration the voc1 EMA
recipes run (``dataset_mode: a2w``, ``package_mode: random_window``, optional CAR past-sample
slice): same constructor keywords, same use of ``np.random.randint`` (so a seeded numpy RNG
yields the same windows as the reference), same integer indexing — checked bit-exactly against
the reference's own output in tests/golden/collater.npz.  Pure numpy / host work: the batch it
returns is what ``TrainStep.step`` copies host->device.
"""
import queue
import threading

import numpy as np
import torch


class SpeechCollater(object):
    """Customized collater for the PyTorch DataLoader in training (reference bin/train.py:865)."""

    def __init__(self, batch_max_steps=20480, hop_size=256, aux_context_window=0, use_noise_input=False,
                 dataset_mode="a2w", use_spk_id=False, use_ph=False, config=None, rng=None):
        """``rng``: object with ``randint(lo, hi)`` drawing the window starts; default = the global ``np.random``
        (as the reference).  A collater used beside a prefetching train loop (dev batches) gets its own
        ``np.random.RandomState`` so that it does not perturb the train windows."""
        assert batch_max_steps % hop_size == 0
        self.rng = np.random if rng is None else rng
        if dataset_mode != "a2w":
            raise NotImplementedError("only dataset_mode 'a2w' (articulatory -> waveform) is on the B200 hot path")
        if use_spk_id or use_ph:
            raise NotImplementedError("speaker / phoneme conditioning is outside the B200 hot path")
        config = config or {}
        if config.get("package_mode", "random_window") != "random_window":
            raise NotImplementedError("only package_mode 'random_window' (the reference default) is implemented")
        if "generator2_params" in config or "generator2_type" in config:
            raise NotImplementedError("two-stage generator cascades are outside the B200 hot path")
        self.batch_max_steps = batch_max_steps
        self.batch_max_frames = batch_max_steps // hop_size
        self.hop_size = hop_size
        self.aux_context_window = aux_context_window
        self.use_noise_input = use_noise_input
        self.dataset_mode = dataset_mode
        gp = config.get("generator_params", {})
        self.use_ar = gp.get("use_ar", False)
        # a2w: the AR context is past WAVEFORM samples (reference :897-905, ar2_len)
        self.ar_len = int(gp.get("ar_input", 512) / gp.get("out_channels", 1)) if self.use_ar else None
        self.start_offset = aux_context_window                                   # :919
        self.end_offset = -(self.batch_max_frames + aux_context_window)          # :920
        self.config = config

    def window_plan(self, audio_len, art_len, start_frame):
        """Integer index plan of one item for a given start frame (reference :1009-1027, :1082-1097)."""
        w0 = start_frame * self.hop_size
        plan = dict(art=(start_frame - self.aux_context_window,
                         start_frame + self.batch_max_frames + self.aux_context_window),
                    wav=(w0, w0 + self.batch_max_steps))
        if self.use_ar:
            lo = w0 - self.ar_len
            plan["ar"] = (max(lo, 0), w0)
            plan["ar_left_zero_pad"] = max(-lo, 0)
        return plan

    def __call__(self, batch):
        """batch: list of dicts with 'audio' (T,) and 'art' (T', C) numpy arrays.
        Returns dict with 'x' = ((B, C, T') float,), 'y' = (B, 1, T) float [, 'ar' = (B, 1, ar_len)]."""
        audios, arts = [], []
        for d in batch:
            audio, art = d["audio"], d["art"]
            art = art[:int(len(audio) / self.hop_size)]                          # :983
            if len(art) + self.end_offset > self.start_offset:                   # :984
                audios.append(audio)
                arts.append(art)
        # one np.random.randint per kept item, in order (reference :1013)
        starts = np.array([self.rng.randint(self.start_offset, len(c) + self.end_offset) for c in arts])
        plans = [self.window_plan(len(a), len(c), int(s)) for a, c, s in zip(audios, arts, starts)]
        audio_batch = np.stack([a[p["wav"][0]:p["wav"][1]] for a, p in zip(audios, plans)], axis=0)
        art_batch = np.stack([c[p["art"][0]:p["art"][1]] for c, p in zip(arts, plans)], axis=0)
        out = {"audio": torch.tensor(audio_batch, dtype=torch.float).unsqueeze(1),          # (B, 1, T)
               "art": torch.tensor(art_batch, dtype=torch.float).transpose(2, 1)}           # (B, C, T')
        out["x"] = (out["art"],)
        out["y"] = out["audio"]
        if self.use_ar:
            ars = []
            for a, p in zip(audios, plans):
                ar = a[p["ar"][0]:p["ar"][1]]
                if p["ar_left_zero_pad"]:
                    ar = np.pad(ar, (p["ar_left_zero_pad"], 0), "constant", constant_values=0)
                ars.append(ar)
            out["ar"] = torch.tensor(np.stack(ars, axis=0), dtype=torch.float).unsqueeze(1)  # (B, 1, ar_len)
        return out


class DeviceWindowCutter(object):
    """The dataset lives in HBM; a batch is cut ON THE DEVICE (reference data path: datasets/audio_mel_dataset.py:305-531
    feeding SpeechCollater, bin/train.py:965-1098).

    All utterances are uploaded once — MNGU0 is ~230 MB of fp32 audio + ~40 MB of features against 180 GB of HBM3e — and
    every step costs one 8-byte-per-item index upload and one gather launch (``artic_cut_windows``) instead of numpy
    slicing, a 628 KB host-to-device copy and a pinning thread.  The window starts are drawn on the host by the
    collater's own RNG calls, in the collater's order, so a seeded run cuts EXACTLY the collater's windows
    (``tests/test_gpu_plugins.py::test_device_window_cutter_matches_collater``)."""

    def __init__(self, items, collater, device):
        from ._lib import require_cuda
        self.c, self.dev = collater, torch.device(device)
        hop = collater.hop_size
        self.art_len = [min(len(d["art"]), int(len(d["audio"]) / hop)) for d in items]            # bin/train.py:983
        self.keep = [n + collater.end_offset > collater.start_offset for n in self.art_len]       # :984
        audio = [np.asarray(d["audio"], dtype=np.float32).reshape(-1) for d in items]
        art = [np.asarray(d["art"], dtype=np.float32)[:n] for d, n in zip(items, self.art_len)]
        self.C = art[0].shape[1]
        a_off = np.cumsum([0] + [len(a) for a in audio])[:-1]
        t_off = np.cumsum([0] + [a.size for a in art])[:-1]
        self.audio = torch.from_numpy(np.concatenate(audio)).to(self.dev)
        self.art = torch.from_numpy(np.concatenate([a.reshape(-1) for a in art])).to(self.dev)
        self.a_off = torch.from_numpy(a_off.astype(np.int64)).to(self.dev)
        self.t_off = torch.from_numpy(t_off.astype(np.int64)).to(self.dev)
        require_cuda(self.audio, "dataset")

    def picks(self, indices):
        """[(utterance, start frame)] of a batch: one ``rng.randint`` per kept item, in order (bin/train.py:1013)."""
        c = self.c
        return [(i, int(c.rng.randint(c.start_offset, self.art_len[i] + c.end_offset))) for i in indices if self.keep[i]]

    def __call__(self, indices):
        """Device batch {'x': ((B, C, T'),), 'y': (B, 1, T), ['ar': (B, 1, ar_len)]} of the utterances ``indices``."""
        from ._lib import call, ptr
        c = self.c
        pk = self.picks(indices)
        B, aux, frames = len(pk), c.aux_context_window, c.batch_max_frames
        pick = torch.tensor(pk, dtype=torch.int32).reshape(B, 2).pin_memory().to(self.dev, non_blocking=True)
        x = torch.empty((B, self.C, frames + 2 * aux), dtype=torch.float32, device=self.dev)
        y = torch.empty((B, 1, c.batch_max_steps), dtype=torch.float32, device=self.dev)
        ar_len = c.ar_len if c.use_ar else 0
        ar = torch.empty((B, 1, ar_len), dtype=torch.float32, device=self.dev) if c.use_ar else None
        call("artic_cut_windows", ptr(self.audio), ptr(self.a_off), ptr(self.art), ptr(self.t_off), ptr(pick), B, self.C,
             frames, aux, c.hop_size, ar_len, ptr(x), ptr(y), ptr(ar))
        out = {"art": x, "audio": y, "x": (x,), "y": y}
        if c.use_ar:
            out["ar"] = ar
        return out


class BatchPrefetcher(object):
    """Assemble (and pin) the next batches on ONE host thread while the GPU runs the current step.

    ``make_batch(job)`` is called for every job in order on the worker thread — a single worker, so a seeded
    ``np.random`` (the collater's window draws) yields exactly the batches of the synchronous loop — and its
    results are handed over through a bounded queue (``depth`` batches ahead).  With ``pin=True`` every tensor of a
    dict result is moved to pinned memory, ready for the non-blocking host->device copy inside ``TrainStep.step``.
    Exceptions of the worker are re-raised by the iterator; ``close()`` (also called when iteration ends or is
    abandoned) stops the worker."""

    _END = object()

    def __init__(self, make_batch, jobs, depth=2, pin=True, device=None):
        """``device``: the rank's CUDA device.  The CUDA current device is per THREAD: without setting it here the
        worker would pin through a fresh context on cuda:0 on every rank."""
        self._q = queue.Queue(maxsize=max(int(depth), 1))
        self._stop = threading.Event()
        self._pin = pin and torch.cuda.is_available()

        def work():
            try:
                if device is not None and torch.cuda.is_available():
                    torch.cuda.set_device(device)
                for job in jobs:
                    if self._stop.is_set():
                        return
                    b = make_batch(job)
                    if self._pin and isinstance(b, dict):
                        b = {k: (v.pin_memory() if torch.is_tensor(v) else
                                 tuple(t.pin_memory() for t in v) if isinstance(v, tuple) else v) for k, v in b.items()}
                    if not self._put((None, b)):
                        return
                self._put((None, self._END))
            except BaseException as ex:                                    # noqa: BLE001 — handed to the consumer
                self._put((ex, None))

        self._t = threading.Thread(target=work, name="artic-batch-prefetch", daemon=True)
        self._t.start()

    def _put(self, item):
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def __iter__(self):
        try:
            while True:
                ex, b = self._q.get()
                if ex is not None:
                    raise ex
                if b is self._END:
                    return
                yield b
        finally:
            self.close()

    def close(self):
        self._stop.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass
        self._t.join(timeout=5.0)


def synthetic_utterances(n_items, n_frames=400, n_feats=13, hop_size=80, seed=0):
    """Synthetic MNGU0-shaped items ({'audio', 'art'}) for smoke runs without a dataset."""
    rng = np.random.RandomState(seed)
    items = []
    for _ in range(n_items):
        t = np.arange(n_frames * hop_size)
        f0 = rng.uniform(80, 300)
        audio = (0.5 * np.sin(2 * np.pi * f0 * t / 16000.0 + rng.uniform(0, 6.28)) + 0.05 * rng.randn(len(t)))
        items.append({"audio": np.clip(audio, -1, 1).astype(np.float32),
                      "art": rng.randn(n_frames, n_feats).astype(np.float32)})
    return items
