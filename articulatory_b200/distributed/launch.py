#!/usr/bin/env python
"""Process spawner with the flags of the reference launcher (articulatory/distributed/launch.py:
--nnodes, --node_rank, --nproc_per_node, --master_addr, --master_port, --use_env, -m, -c,
training_script ...): one process per GPU, rendezvous through MASTER_ADDR / MASTER_PORT /
WORLD_SIZE / RANK / LOCAL_RANK in the environment (``init_method="env://"``, as reference
bin/train.py:1455-1459 expects).  `run.sh` stage 2 calls it as
``python -m articulatory_b200.distributed.launch --nproc_per_node N -c articulatory-train ...``.

Differences: a failing rank terminates its siblings instead of leaving them blocked in a
collective, and the exit status is the first non-zero status.
"""
import argparse
import os
import subprocess
import sys
import time


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Spawn one training process per GPU")
    p.add_argument("--nnodes", type=int, default=1)
    p.add_argument("--node_rank", type=int, default=0)
    p.add_argument("--nproc_per_node", type=int, default=1)
    p.add_argument("--master_addr", default="127.0.0.1", type=str)
    p.add_argument("--master_port", default=29500, type=int)
    p.add_argument("--use_env", default=False, action="store_true",
                   help="pass the local rank only through LOCAL_RANK (no --local_rank argument)")
    p.add_argument("-m", "--module", default=False, action="store_true")
    p.add_argument("-c", "--command", default=False, action="store_true")
    p.add_argument("training_script", type=str)
    p.add_argument("training_script_args", nargs=argparse.REMAINDER)
    return p.parse_args(argv)


def rank_env(args, local_rank, base=None):
    """Environment of one rank (pure function: unit-tested on CPU)."""
    env = dict(os.environ if base is None else base)
    env["MASTER_ADDR"] = args.master_addr
    env["MASTER_PORT"] = str(args.master_port)
    env["WORLD_SIZE"] = str(args.nproc_per_node * args.nnodes)
    env["RANK"] = str(args.nproc_per_node * args.node_rank + local_rank)
    env["LOCAL_RANK"] = str(local_rank)
    if "OMP_NUM_THREADS" not in env and args.nproc_per_node > 1:
        env["OMP_NUM_THREADS"] = "1"
    return env


def rank_cmd(args, local_rank):
    if args.command:
        cmd = [args.training_script]
    else:
        cmd = [sys.executable, "-u"] + (["-m"] if args.module else []) + [args.training_script]
    if not args.use_env:
        cmd.append(f"--local_rank={local_rank}")
    return cmd + list(args.training_script_args)


def main(argv=None):
    args = parse_args(argv)
    procs = [subprocess.Popen(rank_cmd(args, r), env=rank_env(args, r)) for r in range(args.nproc_per_node)]
    status = 0
    alive = list(procs)
    while alive:
        for p in list(alive):
            rc = p.poll()
            if rc is None:
                continue
            alive.remove(p)
            if rc != 0 and status == 0:
                status = rc
                for q in alive:          # do not leave siblings hanging in a collective
                    q.terminate()
        time.sleep(0.05)
    if status != 0:
        raise subprocess.CalledProcessError(returncode=status, cmd=rank_cmd(args, 0))


if __name__ == "__main__":
    main()
