#!/usr/bin/env python
"""Train a HiFi-GAN / HiFi-CAR vocoder on the B200 hot path (reference bin/train.py).

Same flags as ``articulatory-train`` (--train-dumpdir, --dev-dumpdir, --outdir, --config,
--pretrain, --resume, --verbose, --rank, --local_rank), same YAML keys
(``egs/ema/voc1/conf/e2w_hifigan*.yaml`` run unchanged), same plugin lookup by class name,
same checkpoint layout (``{"model": {"generator", "discriminator"}, "optimizer", "scheduler",
"steps", "epochs"}`` with ``torch.optim.Adam`` / ``MultiStepLR`` state dicts, so checkpoints resume across the two
code bases), same schedule gates and logged scalars.  The loop body is the fused
``TrainStep`` (three CUDA graphs per step, no per-step host sync); distributed training — disabled
in the reference (bin/train.py:1790-1801) — is one process per GPU with an NCCL gradient all-reduce.

Data: ``--train-dumpdir`` with ``*-wave.npy`` / ``*-feats.npy`` pairs (the reference's ``format:
npy`` dump layout) or ``*.h5`` with ``wave`` / ``feats`` datasets when h5py is installed;
``--synthetic N`` trains on N generated MNGU0-shaped utterances (no dataset needed).
"""
import argparse
import logging
import os
import sys

import numpy as np
import torch
import yaml

import articulatory_b200
import articulatory_b200.models
from articulatory_b200.data import BatchPrefetcher, DeviceWindowCutter, SpeechCollater, synthetic_utterances
from articulatory_b200.parallel import DataParallel
from articulatory_b200.trainer import LOG_KEYS, TrainStep


def _load_items(dumpdir, config):
    """[{'audio': (T,), 'art': (T', C)}] from a dump directory.  With the recipe's ``data/<stage>/feats.scp`` in the
    working directory this is the reference ``SpeechDataset`` (wave from the dump, articulatory features from the
    scp, bin/train.py:1571-1582); without it, ``*-wave.npy`` / ``*-feats.npy`` (or ``wave`` / ``feats`` of each
    ``*.h5``) pairs of the dump itself."""
    from articulatory_b200.datasets import SpeechDataset, find_files, read_hdf5
    fmt = config.get("format", "hdf5")
    if fmt == "hdf5":
        audio_query, mel_query = "*.h5", "*.h5"
        audio_load_fn, mel_load_fn = (lambda x: read_hdf5(x, "wave")), (lambda x: read_hdf5(x, "feats"))
    elif fmt == "npy":
        audio_query, mel_query = "*-wave.npy", "*-feats.npy"
        audio_load_fn = mel_load_fn = np.load
    else:
        raise ValueError("support only hdf5 or npy format.")
    parts = dumpdir.split("/")
    # reference bin/train.py:1511-1516,1547: utterances not longer than the training window are dropped up front
    mel_thr = None
    if config.get("remove_short_samples", False):
        mel_thr = config["batch_max_steps"] // config["hop_size"] + 2 * config["generator_params"].get("aux_context_window", 0)
    if len(parts) > 1 and os.path.exists(os.path.join("data", parts[1], "feats.scp")):
        ds = SpeechDataset(dumpdir, audio_query=audio_query, mel_query=mel_query, audio_load_fn=audio_load_fn,
                           mel_load_fn=mel_load_fn, dataset_mode=config.get("dataset_mode", "a2w"),
                           mel_length_threshold=mel_thr)
        return [{"audio": np.asarray(it["audio"], dtype=np.float32), "art": np.asarray(it["art"], dtype=np.float32)}
                for it in (ds[i] for i in range(len(ds)))]
    items = []
    for wav, feats in zip(sorted(find_files(dumpdir, audio_query)), sorted(find_files(dumpdir, mel_query))):
        audio, art = np.asarray(audio_load_fn(wav), dtype=np.float32), np.asarray(mel_load_fn(feats), dtype=np.float32)
        if mel_thr is not None and len(art) <= mel_thr:
            continue
        items.append({"audio": audio, "art": art})
    return items


class ScalarWriter(object):
    """The reference logs every averaged scalar to TensorBoard (``tensorboardX.SummaryWriter(outdir)``, bin/train.py:110,
    :746-748).  tensorboardX is an optional dependency: when it is not installed the same (tag, value, step) triples go
    to ``<outdir>/scalars.jsonl``, one JSON object per line."""

    def __init__(self, outdir):
        self.tb, self.path = None, os.path.join(outdir, "scalars.jsonl")
        try:
            from tensorboardX import SummaryWriter
            self.tb = SummaryWriter(outdir)
        except ImportError:
            pass

    def add_scalars(self, values, step):
        import json
        if self.tb is not None:
            for k, v in values.items():
                self.tb.add_scalar(k, v, step)
        else:
            with open(self.path, "a") as f:
                f.write(json.dumps({"step": int(step), **{k: float(v) for k, v in values.items()}}) + "\n")


class Trainer(object):
    """Epoch / step loop, logging and checkpointing around ``TrainStep`` (reference Trainer, bin/train.py:60-780)."""

    def __init__(self, steps, epochs, items, collater, model, step_fn, config, dp, device, dev_items=None,
                 dev_collater=None, device_data=True):
        self.steps, self.epochs = steps, epochs
        # the training set lives in HBM and windows are cut by a kernel (data.DeviceWindowCutter); device_data=False
        # keeps the host path: numpy slicing + pinning on a prefetch thread
        self.cutter = DeviceWindowCutter(items, collater, device) if device_data else None
        self.items, self.collater, self.model, self.ts = items, collater, model, step_fn
        self.dev_collater = dev_collater or collater
        self.config, self.dp, self.device = config, dp, device
        self.dev_items = dev_items or []
        self.best_mel_loss = float("inf")
        self.finish_train = False
        self.writer = ScalarWriter(config["outdir"]) if dp.rank == 0 else None

    @torch.no_grad()
    def save_intermediate_result(self, batch):
        """Reference Trainer._generate_and_save_intermediate_result (bin/train.py:651-744): generate the first dev batch
        with the current generator and write ``predictions/<steps>steps/<idx>_{ref,gen}.wav`` (PCM-16) for the first
        ``num_save_intermediate_results`` items (+ the reference's two-panel ``<idx>.png`` when matplotlib is there)."""
        from articulatory_b200.bin.decode import _write_wav
        n_save = int(self.config.get("num_save_intermediate_results", 4))
        if n_save <= 0 or self.dp.rank != 0:
            return
        self.ts.sync_exchange()
        G = self.model["generator"]
        x, y = batch["x"][0].to(self.device), batch["y"]
        ar = batch["ar"].to(self.device) if "ar" in batch else None
        y_ = G(x, ar=ar).float().cpu()
        dirname = os.path.join(self.config["outdir"], f"predictions/{self.steps}steps")
        os.makedirs(dirname, exist_ok=True)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            plt = None
        for idx, (r, g) in enumerate(zip(y, y_), 1):
            r, g = r.reshape(-1).numpy(), g.reshape(-1).numpy()
            if plt is not None:
                plt.subplot(2, 1, 1)
                plt.plot(r)
                plt.title("groundtruth speech")
                plt.subplot(2, 1, 2)
                plt.plot(g)
                plt.title(f"generated speech @ {self.steps} steps")
                plt.tight_layout()
                plt.savefig(os.path.join(dirname, f"{idx}.png"))
                plt.close()
            _write_wav(os.path.join(dirname, f"{idx}_ref.wav"), r, self.config["sampling_rate"])
            _write_wav(os.path.join(dirname, f"{idx}_gen.wav"), g, self.config["sampling_rate"])
            if idx >= n_save:
                break

    def eval_epoch(self):
        """Reference Trainer._eval_epoch (bin/train.py:605-648): average the eval losses over the dev set, keep the
        checkpoint with the best mel loss (best_mel_ckpt.pkl + best_mel_step.txt).  Returns the averages."""
        bs = self.config["batch_size"]
        n = 0
        for lo in range(0, len(self.dev_items) - bs + 1, bs):
            batch = self.dev_collater(self.dev_items[lo:lo + bs])
            if batch["y"].shape[0] != bs:
                continue
            self.ts.eval_step(batch["x"][0], batch["y"], batch.get("ar"))
            n += 1
            if n == 1:                                                     # reference :617-618
                self.save_intermediate_result(batch)
        if n == 0:
            return {}
        logs = self.ts.read_eval_logs(n)
        if self.dp.rank == 0:
            logging.info(f"(Steps: {self.steps}) Finished evaluation ({n} steps per epoch).")
            for k, v in logs.items():
                logging.info(f"(Steps: {self.steps}) {k} = {v:.4f}.")
            self.writer.add_scalars(logs, self.steps)                      # reference :640
            if logs["eval/mel_loss"] < self.best_mel_loss:
                with open(os.path.join(self.config["outdir"], "best_mel_step.txt"), "w+") as ouf:
                    ouf.write("%d\n" % self.steps)
                self.save_checkpoint(os.path.join(self.config["outdir"], "best_mel_ckpt.pkl"))
                self.best_mel_loss = logs["eval/mel_loss"]
        return logs

    def save_checkpoint(self, path):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save({"model": {"generator": self.model["generator"].state_dict(),
                              "discriminator": self.model["discriminator"].state_dict()},
                    "optimizer": {"generator": self.ts.optG.state_dict(), "discriminator": self.ts.optD.state_dict()},
                    "scheduler": {"generator": self.ts.optG.scheduler_state_dict(),
                                  "discriminator": self.ts.optD.scheduler_state_dict()},
                    "steps": self.steps, "epochs": self.epochs}, path)

    def load_checkpoint(self, path, load_only_params=False):
        sd = torch.load(path, map_location="cpu", weights_only=False)
        self.model["generator"].load_state_dict(sd["model"]["generator"])
        self.model["discriminator"].load_state_dict(sd["model"]["discriminator"])
        for m in self.model.values():
            m.mark_weights_dirty()
        if not load_only_params:
            self.steps, self.epochs = sd["steps"], sd["epochs"]
            self.ts.steps = self.steps
            sch = sd.get("scheduler", {})
            self.ts.optG.load_state_dict(sd["optimizer"]["generator"], sch.get("generator"))
            self.ts.optD.load_state_dict(sd["optimizer"]["discriminator"], sch.get("discriminator"))

    def run(self):
        bs = self.config["batch_size"]
        while not self.finish_train:
            idx = self.dp.sampler_indices(len(self.items), self.epochs, shuffle=True)
            groups = [idx[lo:lo + bs] for lo in range(0, len(idx) - bs + 1, bs)]
            if not groups:
                raise ValueError(f"{len(idx)} training utterances on this rank cannot fill one batch of {bs}")
            stepped = False
            # device path: one index upload + one gather launch per batch; host path: window cutting + pinning of the
            # next two batches on a host thread under the current step
            batches = (self.cutter(g) for g in groups) if self.cutter is not None else \
                BatchPrefetcher(lambda g: self.collater([self.items[i] for i in g]), groups, depth=2, device=self.device)
            for batch in batches:
                if batch["y"].shape[0] != bs:
                    continue          # an utterance shorter than the window was dropped: keep shapes static
                stepped = True
                self.ts.step(batch["x"][0], batch["y"], batch.get("ar"))
                self.steps += 1
                if self.steps % self.config["log_interval_steps"] == 0:
                    n = self.config["log_interval_steps"]
                    self.ts.sync_exchange()
                    vals = self.dp.mean_scalars(self.ts.running.clone()).cpu().tolist()
                    self.ts.running.zero_()
                    if self.dp.rank == 0:
                        for k, v in zip(LOG_KEYS, vals):
                            logging.info(f"(Steps: {self.steps}) {k} = {v / n:.4f}.")
                        self.writer.add_scalars({k: v / n for k, v in zip(LOG_KEYS, vals)}, self.steps)   # reference :763-773
                if self.dev_items and self.steps % self.config.get("eval_interval_steps", 1000) == 0:
                    self.eval_epoch()                                          # reference :766-768
                if self.steps % self.config["save_interval_steps"] == 0 and self.dp.rank == 0:
                    self.save_checkpoint(os.path.join(self.config["outdir"], f"checkpoint-{self.steps}steps.pkl"))
                if self.steps >= self.config["train_max_steps"]:
                    self.finish_train = True
                    break
            if not stepped:
                raise ValueError("every batch of the epoch lost an utterance shorter than the training window; set "
                                 "remove_short_samples: true (the yaml default) or lower batch_max_steps")
            self.epochs += 1


def main(argv=None):
    parser = argparse.ArgumentParser(description="Train HiFi-GAN / HiFi-CAR (See detail in articulatory_b200/bin/train.py).")
    for name in ("--train-wav-scp", "--train-feats-scp", "--train-segments", "--train-dumpdir", "--train-dumpdirs",
                 "--dev-wav-scp", "--dev-feats-scp", "--dev-segments", "--dev-dumpdir", "--dev-dumpdirs"):
        parser.add_argument(name, default=None, type=str)
    parser.add_argument("--outdir", type=str, required=True, help="directory to save checkpoints.")
    parser.add_argument("--config", type=str, required=True, help="yaml format configuration file.")
    parser.add_argument("--pretrain", default="", type=str, help='checkpoint file path to load pretrained params. (default="")')
    parser.add_argument("--pretrain2", default="", type=str)
    parser.add_argument("--resume", default="", type=str, help='checkpoint file path to resume training. (default="")')
    parser.add_argument("--verbose", type=int, default=1)
    parser.add_argument("--rank", "--local_rank", default=0, type=int)
    parser.add_argument("--synthetic", type=int, default=0, help="train on N synthetic utterances (B200 extension)")
    parser.add_argument("--host-data", action="store_true",
                        help="cut the training windows on the host (numpy + pinned prefetch thread) instead of on the device "
                             "from a dataset resident in HBM (B200 extension)")
    parser.add_argument("--max-steps", type=int, default=0,
                        help="stop after this many steps instead of the yaml's train_max_steps (B200 extension: smoke runs "
                             "of an unchanged recipe yaml)")
    parser.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"],
                        help="bf16x3 (default): fp32 storage, error-compensated split-bf16 tcgen05 contraction — the mode "
                             "that meets the 1e-3 parity gate against the fp32 reference; bf16: bf16 storage, plain "
                             "tcgen05 (fastest, ~1e-2 waveform error); fp32: CUDA-core kernels (debug)")
    args = parser.parse_args(argv)

    if not torch.cuda.is_available():
        raise RuntimeError("articulatory_b200 trains on a CUDA device (sm_100a); there is no CPU path")
    dp_probe = int(os.environ.get("LOCAL_RANK", args.rank))
    torch.cuda.set_device(dp_probe)
    device = torch.device("cuda", dp_probe)
    dp = DataParallel(device=device)
    args.distributed = dp.world > 1
    if dp.rank != 0:
        sys.stdout = open(os.devnull, "w")                                   # reference :1462-1463
    logging.basicConfig(level=logging.INFO if args.verbose > 0 else logging.WARN, stream=sys.stdout,
                        format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s")
    os.makedirs(args.outdir, exist_ok=True)
    with open(args.config) as f:
        config = yaml.load(f, Loader=yaml.Loader)
    config.update(vars(args))
    if args.max_steps > 0:
        config["train_max_steps"] = args.max_steps
    config["version"] = articulatory_b200.__version__
    if dp.rank == 0:
        with open(os.path.join(args.outdir, "config.yml"), "w") as f:
            yaml.dump(config, f, Dumper=yaml.Dumper)
    if args.pretrain2 or "generator2_type" in config:
        raise NotImplementedError("two-stage generator cascades are outside the B200 hot path")

    if args.synthetic:
        hop = config["hop_size"]
        frames = 4 * config["batch_max_steps"] // hop
        n_feats = config["generator_params"]["in_channels"] - (config["generator_params"].get("ar_output", 0)
                                                               if config["generator_params"].get("use_ar") else 0)
        items = synthetic_utterances(args.synthetic, frames, n_feats, hop, seed=dp.rank)
    elif args.train_dumpdir is not None:
        items = _load_items(args.train_dumpdir, config)
    else:
        raise ValueError("Please specify --train-dumpdir or --synthetic.")
    logging.info(f"The number of training files = {len(items)}.")
    dev_items = _load_items(args.dev_dumpdir, config) if args.dev_dumpdir is not None else \
        (synthetic_utterances(max(config["batch_size"], 2), frames, n_feats, hop, seed=10_000 + dp.rank) if args.synthetic else [])
    logging.info(f"The number of development files = {len(dev_items)}.")
    collater = SpeechCollater(batch_max_steps=config["batch_max_steps"], hop_size=config["hop_size"],
                              aux_context_window=config["generator_params"].get("aux_context_window", 0),
                              dataset_mode=config.get("dataset_mode", "a2w"), config=config)

    # plugin lookup by class name (reference :1649-1662)
    generator_class = getattr(articulatory_b200.models, config.get("generator_type", "HiFiGANGenerator"))
    discriminator_class = getattr(articulatory_b200.models,
                                  config.get("discriminator_type", "HiFiGANMultiScaleMultiPeriodDiscriminator"))
    model = {"generator": generator_class(**config["generator_params"], precision=args.precision).to(device),
             "discriminator": discriminator_class(**config["discriminator_params"], precision=args.precision).to(device)}
    dp.broadcast_parameters(model["generator"], model["discriminator"])
    ts = TrainStep(model["generator"], model["discriminator"], config, device, world_size=dp.world,
                   all_reduce=dp.all_reduce if dp.world > 1 else None,
                   grad_wire=dp.wire_of if dp.world > 1 else None)
    dev_collater = SpeechCollater(batch_max_steps=config["batch_max_steps"], hop_size=config["hop_size"],
                                  aux_context_window=config["generator_params"].get("aux_context_window", 0),
                                  dataset_mode=config.get("dataset_mode", "a2w"), config=config,
                                  rng=np.random.RandomState(777 + dp.rank))   # own stream: the prefetch thread owns np.random
    trainer = Trainer(0, 0, items, collater, model, ts, config, dp, device, dev_items=dev_items, dev_collater=dev_collater,
                      device_data=not args.host_data)
    if args.pretrain:
        trainer.load_checkpoint(args.pretrain, load_only_params=True)
    if args.resume:
        trainer.load_checkpoint(args.resume)
    try:
        trainer.run()
        if args.max_steps > 0 and dev_items and trainer.steps % config.get("eval_interval_steps", 1000) != 0:
            trainer.eval_epoch()          # a shortened run still ends with one evaluation pass (+ intermediate results)
    finally:
        if dp.rank == 0:
            trainer.save_checkpoint(os.path.join(config["outdir"], f"checkpoint-{trainer.steps}steps.pkl"))
            logging.info(f"Successfully saved checkpoint @ {trainer.steps}steps.")
        dp.close()


if __name__ == "__main__":
    main()
