"""AMP discriminator-reward path: input assembly and reward epilogue on the GPU kernels, the
dense MLP in between stays torch/cuBLAS (out of scope: dense GEMMs).

Reference: rsl_rl/rsl_rl/algorithms/amp_discriminator.py (DISC) `AMPDiscriminator`,
`Normalizer` / `RunningMeanStd` rsl_rl/rsl_rl/utils/utils.py:79-130 (UT), terminal patch
rsl_rl/rsl_rl/runners/hybrid_runner.py:191-192.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L


class Normalizer:
    """UT:113-130 with the statistics resident on the device (float64) and the batch moments
    computed by hl_column_moments instead of a host round trip (hybrid_ppo.py:279-281)."""

    def __init__(self, input_dim, epsilon=1e-4, clip_obs=10.0, device="cuda:0", process_group=None):
        self.device = torch.device(device)
        self.dim = int(input_dim)
        self.mean = torch.zeros(self.dim, dtype=torch.float64, device=self.device)
        self.var = torch.ones(self.dim, dtype=torch.float64, device=self.device)
        self.count = 1e-4                                    # RunningMeanStd epsilon default, UT:87
        self.epsilon, self.clip_obs = epsilon, clip_obs
        self.process_group = process_group
        self._ws = torch.zeros(int(L.lib.hl_moments_workspace_bytes(self.dim)), dtype=torch.uint8, device=self.device)
        self._mv = torch.zeros(2 * self.dim, dtype=torch.float64, device=self.device)

    def update(self, arr):
        """UT:90-110 (Chan parallel-variance merge) for a device batch (M, dim)."""
        arr = arr.detach().to(self.device, torch.float32).contiguous()
        m = arr.shape[0]
        L.check(L.lib.hl_column_moments(L.ptr(arr), m, self.dim, L.ptr(self._mv), L.ptr(self._ws), L.stream()))
        bm, bv, bc = self._mv[:self.dim], self._mv[self.dim:], float(m)
        if self.process_group is not None:
            bm, bv, bc = allreduce_moments(bm, bv, bc, self.process_group)
        self.update_from_moments(bm, bv, bc)

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + delta.square() * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot

    def mean_std_f32(self):
        """What normalize_torch builds on every call (UT:125-128): fp32(mean), sqrt(fp32(var+eps))."""
        return self.mean.to(torch.float32), torch.sqrt((self.var + self.epsilon).to(torch.float32))

    def normalize_torch(self, input, device=None):
        mean, std = self.mean_std_f32()
        return torch.clamp((input - mean) / std, -self.clip_obs, self.clip_obs)


def allreduce_moments(mean, var, count, group):
    """Merge per-rank batch moments into global-batch moments (exact Chan merge via sums)."""
    import torch.distributed as dist
    pack = torch.cat([mean * count, (var + mean.square()) * count, mean.new_tensor([count])])
    dist.all_reduce(pack, group=group if group != "world" else None)
    d = mean.numel()
    tot = float(pack[-1].item())
    gm = pack[:d] / tot
    gv = pack[d:2 * d] / tot - gm.square()
    return gm, gv.clamp_min(0.0), tot


class AMPDiscriminator(nn.Module):
    """DISC:9-71: MLP input_dim -> hidden -> 1; `predict_amp_reward` runs the fused assembly and
    epilogue kernels around the (cuBLAS) MLP."""

    def __init__(self, input_dim, amp_reward_coef, hidden_layer_sizes, device, task_reward_lerp=0.0):
        super().__init__()
        self.device = device
        self.input_dim = input_dim
        self.amp_reward_coef = amp_reward_coef
        layers, cur = [], input_dim
        for h in hidden_layer_sizes:
            layers += [nn.Linear(cur, h), nn.ReLU()]
            cur = h
        self.trunk = nn.Sequential(*layers).to(device)
        self.amp_linear = nn.Linear(hidden_layer_sizes[-1], 1).to(device)
        self.trunk.train()
        self.amp_linear.train()
        self.task_reward_lerp = task_reward_lerp

    def forward(self, x):
        return self.amp_linear(self.trunk(x))

    def compute_grad_pen(self, expert_state, expert_next_state, lambda_=10):
        data = torch.cat([expert_state, expert_next_state], dim=-1)
        data.requires_grad = True
        disc = self.amp_linear(self.trunk(data))
        grad = torch.autograd.grad(outputs=disc, inputs=data, grad_outputs=torch.ones_like(disc), create_graph=True,
                                   retain_graph=True, only_inputs=True)[0]
        return lambda_ * grad.norm(2, dim=1).pow(2).mean()

    def assemble_input(self, state, next_state, normalizer=None, reset_env_ids=None, terminal_states=None,
                       n_reset_dev=None, return_patched=False):
        """x = cat(normalise(state), normalise(next_state')) (N,60) with next_state' rows of the
        reset envs replaced by their terminal AMP states (hybrid_runner.py:191-192)."""
        n = state.shape[0]
        # the kernel is written for the reference's AMP observation (LR:416: 30 columns, 60-wide discriminator input)
        if state.dim() != 2 or state.shape[1] != 30 or tuple(next_state.shape) != tuple(state.shape):
            raise ValueError(f"assemble_input: state / next_state must be (N, 30), got {tuple(state.shape)} / {tuple(next_state.shape)}")
        if self.input_dim != 60:
            raise ValueError(f"assemble_input: discriminator input_dim must be 60, is {self.input_dim}")
        if state.dtype != torch.float32 or next_state.dtype != torch.float32 or not state.is_cuda or not next_state.is_cuda:
            raise ValueError("assemble_input: float32 CUDA tensors required")
        if normalizer is not None and int(getattr(normalizer, "dim", 30)) != 30:
            raise ValueError("assemble_input: the normaliser must have 30 columns")
        if terminal_states is not None and (terminal_states.dim() != 2 or terminal_states.shape[1] != 30):
            raise ValueError("assemble_input: terminal_states must be (n_reset, 30)")
        state = state.contiguous()
        next_state = next_state.contiguous()
        x = torch.empty(n, 60, device=state.device)
        mean = std = None
        if normalizer is not None:
            mean, std = normalizer.mean_std_f32()
        patched = torch.empty_like(next_state) if return_patched else None
        if reset_env_ids is not None and n_reset_dev is None:
            n_reset_dev = torch.tensor([reset_env_ids.numel()], dtype=torch.int32, device=state.device)
        L.check(L.lib.hl_amp_disc_input(L.ptr(state), L.ptr(next_state), L.ptr(mean), L.ptr(std),
                                        float(normalizer.clip_obs) if normalizer is not None else 0.0,
                                        L.ptr(reset_env_ids), L.ptr(n_reset_dev), L.ptr(terminal_states), L.ptr(patched),
                                        L.ptr(x), n, L.stream()))
        return (x, patched) if return_patched else x

    def predict_amp_reward(self, state, next_state, task_reward, normalizer=None):
        """DISC:55-68 -> (reward (N,), d (N,1))."""
        with torch.no_grad():
            self.eval()
            x = self.assemble_input(state, next_state, normalizer)
            d = self.amp_linear(self.trunk(x))
            reward = torch.empty(state.shape[0], device=state.device)
            L.check(L.lib.hl_amp_reward(L.ptr(d.contiguous()), L.ptr(task_reward.contiguous()), float(self.amp_reward_coef),
                                        float(self.task_reward_lerp), L.ptr(reward), state.shape[0], L.stream()))
            self.train()
        return reward, d
