"""Env-sharded data parallelism: one process per GPU, rank r owns envs
[r*N/W, (r+1)*N/W) of ONE global env set; env state never crosses GPUs.  The reference has no
distributed code at all; these are the only exchange steps the sharded path needs:

  * flat gradient all-reduce (mean) per optimiser step   -- after loss.backward(),
    rsl_rl/rsl_rl/algorithms/him_ppo.py:182, hybrid_ppo.py:271, modules/him_estimator.py:112
  * advantage moments (sum, sum^2, count), 3 float64      -- him_rollout_storage.py:126-127
  * AMP normaliser batch moments                          -- hybrid_ppo.py:279-281 (amp_discriminator.allreduce_moments)
  * the adaptive-KL scalar so every rank takes the same LR branch -- him_ppo.py:144-156

All messages are <= 4.5 MB => latency-bound: one NCCL call per optimiser step on one flat buffer.
Backend: NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests.
"""
import datetime
import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns
    (rank, world, local_rank); world == 1 without WORLD_SIZE (no process group created)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        # every message of this library is <= 4.5 MB and overlaps the env kernels, so what matters is the latency of a
        # message while it shares the SMs.  Measured on 8 GPUs (tools/scale_probe.py, profiles/r2_scale_probe_n8.json;
        # 40 messages per 3.60 ms rollout): NCCL_MAX_CTAS=8 with the tuner's protocol 2.73 ms alone / +0.97 ms overlapped;
        # 2 CTAs 6.56 / +4.9 ms; the LL protocol on up to 32 CTAs 1.70 ms alone / +0.15 ms overlapped.
        os.environ.setdefault("NCCL_MAX_CTAS", "32")
        os.environ.setdefault("NCCL_PROTO", "LL")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        # short collective timeout: a rank that falls out of step must fail fast, not hold a GPU box
        to = datetime.timedelta(seconds=int(os.environ.get("HL_DIST_TIMEOUT_S", "120")))
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, timeout=to, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world, timeout=to)
    return rank, world, local


def shard_range(num_envs_global: int, rank: int, world: int):
    """Contiguous shard [lo, hi) by GLOBAL env id.  Env ids carry meaning in the reference
    (high-speed command envs `env_ids < 0.2 N` legged_robot.py:649-651, stumble slices :1597-1598,
    terrain columns :1234), so shards are cut on the global numbering and the kernels receive
    `env_id_offset`."""
    if num_envs_global % world:
        raise ValueError(f"num_envs {num_envs_global} must divide evenly over {world} ranks")
    per = num_envs_global // world
    return rank * per, (rank + 1) * per


class FlatGradAllReducer:
    """Keeps every parameter's .grad as a view into ONE flat buffer, so the all-reduce after
    backward() is a single collective (in place, no packing kernels)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def zero_grad(self):
        self.flat.zero_()

    def reattach(self):
        """optimizer.zero_grad(set_to_none=True) drops the views; call this instead of it."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + p.numel()].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce_mean(self, async_op: bool = False):
        """grads <- mean over ranks (equal shard sizes => the global-batch gradient)."""
        if self.world == 1:
            return None
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=False)
        self.flat.div_(self.world)
        return work

    def clip_grad_norm_(self, max_norm: float):
        """nn.utils.clip_grad_norm_ on the already-reduced flat buffer (him_ppo.py:183)."""
        total = torch.linalg.vector_norm(self.flat)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(coef)
        return total


def all_reduce_scalar_mean(x: torch.Tensor, group=None) -> torch.Tensor:
    """kl_mean agreement (him_ppo.py:144-156): every rank must take the same LR branch."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    y = x.clone()
    dist.all_reduce(y, group=group)
    return y / dist.get_world_size(group)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t
