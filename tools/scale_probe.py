"""Attribution probe for the N>1 bench: the same rollout graph with the gradient messages arranged five ways.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_probe.py

Prints one JSON line (rank 0): ms per rollout, max over ranks, for
  env          rollout only (moments all-reduce inside)
  comm         the 40 messages only, back to back on one stream
  fork_step    message pair k forked behind env-step k's fused kernel 
  fork_start   all 40 messages forked at the start of the rollout (bench.py's arrangement)
  serial       rollout, then the 40 messages
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B                                    # noqa: E402
from isaacgymloco_b200 import dist as D              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--rollout", type=int, default=24)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    B.bind_to_gpu_numa(local)
    wl = B.Workload(args.envs, args.rollout, rank, world, device)
    stat_group = dist.new_group(backend="nccl")
    grad_group = dist.new_group(backend="nccl")
    wl.storage.process_group = stat_group
    comm = torch.cuda.Stream()
    fork = torch.cuda.Event()
    bufs = [torch.randn(B.EST_PARAMS, device=device), torch.randn(B.AC_PARAMS + 1, device=device)]

    def pair():
        dist.all_reduce(bufs[0], op=dist.ReduceOp.AVG, group=grad_group)
        dist.all_reduce(bufs[1], op=dist.ReduceOp.AVG, group=grad_group)

    def msgs(k0, k1):
        fork.record()
        comm.wait_event(fork)
        with torch.cuda.stream(comm):
            for _ in range(k0, k1):
                pair()

    def arrangement(name):
        cur = torch.cuda.current_stream()
        wl.exchange_cb = None
        if name == "env":
            wl.rollout()
        elif name == "comm":
            msgs(0, 20)
            cur.wait_stream(comm)
        elif name == "fork_step":
            wl.exchange_cb = lambda k: msgs(k, k + 1) if k < 20 else None
            wl.rollout()
            cur.wait_stream(comm)
        elif name == "fork_start":
            msgs(0, 20)
            wl.rollout()
            cur.wait_stream(comm)
        elif name == "serial":
            wl.rollout()
            msgs(0, 20)
            cur.wait_stream(comm)
        wl.exchange_cb = None

    for _ in range(3):
        arrangement("fork_step")
    torch.cuda.synchronize()
    out = {"world": world, "tag": args.tag, "envs": args.envs,
           "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
    for name in ("env", "comm", "fork_step", "fork_start", "serial"):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            arrangement(name)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                arrangement(name)
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(3):
            g.replay()
        ms = B.timed(g.replay, args.steps, True) / args.steps
        out[name] = round(ms, 4)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)          # graphs holding captured NCCL kernels make destroy_process_group() wait for its timeout: leave at once


if __name__ == "__main__":
    main()
