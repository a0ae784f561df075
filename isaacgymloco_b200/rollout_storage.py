"""HIMRolloutStorage with the GAE/returns pass on the GPU kernels.

Reference: rsl_rl/rsl_rl/storage/him_rollout_storage.py:36-177 (`HIMRolloutStorage`); the
`compute_returns` body is byte-identical in amp_rollout_storage.py:141-155 (`RolloutStorage`), so
`RolloutStorage` is exported as an alias.  Same constructor, field names, shapes and dtypes
(time-major (T,N,.) buffers, uint8 dones); `compute_returns(last_values, gamma, lam)` has the same
signature and side effects (fills `returns`, rebinds `advantages`).
"""
import torch

from . import _lib as L


class HIMRolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.next_critic_observations = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape,
                 device="cuda:0", shard_statistics=False, process_group=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HIMRolloutStorage (B200) runs on CUDA only (no CPU fallback)")
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        t, n, dev = num_transitions_per_env, num_envs, self.device
        z = lambda *s: torch.zeros(t, n, *s, device=dev)
        self.observations = z(*obs_shape)
        if privileged_obs_shape[0] is not None:
            self.privileged_observations = z(*privileged_obs_shape)
            self.next_privileged_observations = z(*privileged_obs_shape)
        else:
            self.privileged_observations = None
            self.next_privileged_observations = None
        self.rewards = z(1)
        self.actions = z(*actions_shape)
        self.dones = torch.zeros(t, n, 1, device=dev, dtype=torch.uint8)
        self.actions_log_prob, self.values, self.returns, self.advantages = z(1), z(1), z(1), z(1)
        self.mu, self.sigma = z(*actions_shape), z(*actions_shape)
        self.num_transitions_per_env, self.num_envs = t, n
        self.step = 0
        # (sum adv, sum adv^2, count) in float64; all-reduced over `process_group` when the envs
        # are sharded so every rank normalises with the global-batch statistics
        self._moments = torch.zeros(3, dtype=torch.float64, device=dev)
        self.shard_statistics = bool(shard_statistics)
        self.process_group = process_group

    def add_transitions(self, transition):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(transition.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[s].copy_(transition.critic_observations)
        if self.next_privileged_observations is not None:
            self.next_privileged_observations[s].copy_(transition.next_critic_observations)
        self.actions[s].copy_(transition.actions)
        self.rewards[s].copy_(transition.rewards.view(-1, 1))
        self.dones[s].copy_(transition.dones.view(-1, 1))
        self.values[s].copy_(transition.values)
        self.actions_log_prob[s].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(transition.action_mean)
        self.sigma[s].copy_(transition.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """him_rollout_storage.py:113-127: reverse-time GAE scan, then global advantage
        normalisation with the unbiased std."""
        t, n = self.num_transitions_per_env, self.num_envs
        last_values = last_values.detach().to(self.device, torch.float32).contiguous()
        if self.advantages.data_ptr() == self.returns.data_ptr() or not self.advantages.is_contiguous():
            self.advantages = torch.empty_like(self.returns)
        self.gae_scan(last_values, gamma, lam)
        self.normalize_advantages()

    def gae_scan(self, last_values, gamma, lam):
        t, n = self.num_transitions_per_env, self.num_envs
        self._moments.zero_()
        L.check(L.lib.hl_gae_scan(L.ptr(self.rewards), L.ptr(self.values), L.ptr(self.dones), L.ptr(last_values),
                                  L.ptr(self.returns), L.ptr(self.advantages), L.ptr(self._moments), t, n,
                                  float(gamma), float(lam), L.stream()))

    def normalize_advantages(self):
        t, n = self.num_transitions_per_env, self.num_envs
        if self.shard_statistics and torch.distributed.is_initialized():
            torch.distributed.all_reduce(self._moments, group=self.process_group)
        L.check(L.lib.hl_adv_normalize(L.ptr(self.advantages), L.ptr(self._moments), t * n, L.stream()))

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat = done.permute(1, 0, 2).reshape(-1, 1)
        idx = torch.cat((flat.new_tensor([-1], dtype=torch.int64), flat.nonzero(as_tuple=False)[:, 0]))
        lengths = idx[1:] - idx[:-1]
        return lengths.float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """him_rollout_storage.py:137-177 (torch index gathers; a fused multi-tensor gather is a
        "next" row)."""
        batch = self.num_envs * self.num_transitions_per_env
        mb = batch // num_mini_batches
        indices = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        flat = lambda x: x.flatten(0, 1)
        obs = flat(self.observations)
        critic = flat(self.privileged_observations) if self.privileged_observations is not None else obs
        next_critic = flat(self.next_privileged_observations) if self.next_privileged_observations is not None else obs
        acts, vals, rets = flat(self.actions), flat(self.values), flat(self.returns)
        logp, adv, mu, sigma = flat(self.actions_log_prob), flat(self.advantages), flat(self.mu), flat(self.sigma)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                ids = indices[i * mb:(i + 1) * mb]
                yield (obs[ids], critic[ids], acts[ids], next_critic[ids], vals[ids], adv[ids], rets[ids], logp[ids],
                       mu[ids], sigma[ids])


RolloutStorage = HIMRolloutStorage
