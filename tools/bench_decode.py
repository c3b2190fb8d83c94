#!/usr/bin/env python
"""BASELINE configs[2]: e2w_hifigan_car.yaml HiFi-CAR causal-AR generator inference, batch 32 on one
B200: 32 utterances x 600 frames (48 000 samples each), lock-step chunks of 25 frames (2000 samples),
24 sequential chunks, weight norm removed.  Prints one JSON line (audio-samples/s), next to the
reference algorithm (oracle ar_loop, CPU, one utterance) when --cpu is given."""
import argparse
import json
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    from articulatory_b200 import models as M
    from articulatory_b200.decode import BatchedARDecoder
    from oracle import torch_oracle as O
    gp = dict(O.E2W_GENERATOR_PARAMS, final_scale=80, extra_art=False)      # e2w_hifigan_car.yaml:35-58
    cfg = {"generator_params": gp, "batch_max_steps": 2000, "hop_size": 80, "sampling_rate": 16000}
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = M.HiFiGANGenerator(**gp, precision=args.precision)
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    G.remove_weight_norm()
    G = G.eval().to(dev)
    g = torch.Generator().manual_seed(1234)
    feats = [torch.randn(args.frames, 13, generator=g) for _ in range(args.batch)]
    dec = BatchedARDecoder(G, cfg)
    pinned = [f.pin_memory() for f in feats]
    dec.decode(pinned)                                  # captures the chunk graph
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.iters):
        outs = dec.decode(pinned)
        host = [o.cpu() for o in outs]                  # D2H of the waveforms (end to end)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    n = args.batch * args.frames * 80
    line = {"metric": "audio-samples/sec, e2w_hifigan_car chunked-AR inference", "value": n / (ms / 1e3),
            "unit": "audio-samples/s", "ms_per_batch": ms, "wall_ms_per_batch": 1e3 * (time.perf_counter() - t0) / args.iters,
            "rtf": (ms / 1e3) / (n / 16000.0), "dtype": args.precision,
            "config": {"workload": f"{args.batch} utterances x {args.frames} frames, chunks of 25 frames, "
                                   f"{-(-args.frames // 25)} sequential chunks, CUDA graph per chunk, host in / host out"}}
    if args.cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        x = feats[0][:200]
        t0 = time.perf_counter()
        O.ar_loop(gsd, O.E2W_GENERATOR_PARAMS, x, 2000, 80)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 200 * 80 / dt, "unit": "audio-samples/s", "cores": os.cpu_count(),
                                "kind": "port", "sample": "1 utterance x 200 frames, oracle ar_loop (fp32)"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
