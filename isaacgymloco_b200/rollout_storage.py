"""HIMRolloutStorage with the GAE/returns pass on the GPU kernels.

Reference: rsl_rl/rsl_rl/storage/him_rollout_storage.py:36-177 (`HIMRolloutStorage`); the
`compute_returns` body is byte-identical in amp_rollout_storage.py:141-155 (`RolloutStorage`), so
that method serves both; the AMP `RolloutStorage` class itself (history / wm_feature buffers, recurrent
minibatches) is not part of the path and is not provided.  Same constructor, field names, shapes and dtypes
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
                 device="cuda:0", shard_statistics=False, process_group=None, env_writes_slots=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HIMRolloutStorage (B200) runs on CUDA only (no CPU fallback)")
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        t, n, dev = num_transitions_per_env, num_envs, self.device
        z = lambda *s: torch.zeros(t, n, *s, device=dev)
        # env_writes_slots: one extra slot so FusedLeggedRobot.bind_rollout() can make the env write
        # its observations straight into slot step+1 (slot T carries over into slot 0 at clear())
        self.env_writes_slots = bool(env_writes_slots)
        extra = 1 if self.env_writes_slots else 0
        self._obs_all = torch.zeros(t + extra, n, *obs_shape, device=dev)
        self.observations = self._obs_all[:t]
        if privileged_obs_shape[0] is not None:
            self._priv_all = torch.zeros(t + extra, n, *privileged_obs_shape, device=dev)
            self.privileged_observations = self._priv_all[:t]
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

    # ------------------------------------------------------------------ recording one env step
    def _rows(self, x, width, dtype=torch.float32):
        """(N,width) contiguous device rows of `x` (no copy when it already is)."""
        x = x.detach()
        if dtype == torch.uint8:
            if x.dtype == torch.bool:
                x = x.view(torch.uint8) if x.is_contiguous() else x.to(torch.uint8)
            elif x.dtype != torch.uint8:
                x = (x != 0).to(torch.uint8)
        elif x.dtype != dtype:
            x = x.to(dtype)
        if x.device != self.device:
            x = x.to(self.device)
        x = x.reshape(self.num_envs, width)
        return x if x.is_contiguous() else x.contiguous()

    def _record(self, tr, rewards, dones, time_outs, next_critic, term_ids, term_rows, gamma, term_count=None,
                assume_sorted=True):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s, n = self.step, self.num_envs
        od = self.observations.shape[-1]
        ad = self.actions.shape[-1]
        has_priv = self.privileged_observations is not None
        pd = self.privileged_observations.shape[-1] if has_priv else 1
        keep = []   # sources must outlive the enqueue

        def src(x, width, dtype=torch.float32):
            r = self._rows(x, width, dtype)
            keep.append(r)
            return r

        def pair(x, slot, width, dtype=torch.float32):
            """(src_ptr, dst_ptr); an in-place source (the env wrote the slot already) is skipped."""
            if x is None or slot is None:
                return None, None
            r = src(x, width, dtype)
            if r.data_ptr() == slot.data_ptr():
                return None, None
            return L.ptr(r), L.ptr(slot)

        t = L.HlTransition()
        t.struct_bytes = L.ctypes.sizeof(L.HlTransition)
        t.obs_dim, t.priv_dim, t.act_dim, t.gamma = od, pd, ad, float(gamma)
        t.obs, t.obs_out = pair(tr.observations, self.observations[s], od)
        if has_priv:
            t.critic_obs, t.critic_out = pair(tr.critic_observations, self.privileged_observations[s], pd)
            t.next_critic_obs, t.next_critic_out = pair(next_critic, self.next_privileged_observations[s], pd)
            if term_ids is not None and t.next_critic_out:
                ids = term_ids.detach().to(self.device, torch.int64).flatten().contiguous()
                rows = term_rows.detach().to(self.device, torch.float32).reshape(-1, pd).contiguous()
                if not assume_sorted:      # reset_buf.nonzero() is ascending; other id lists are sorted here (no host sync)
                    ids, order = torch.sort(ids, stable=True)
                    rows = rows[order]
                if term_count is not None:  # device-side count of over-allocated id/row buffers (sync-free env.step)
                    cnt = term_count.detach().to(self.device, torch.int32).reshape(1)
                else:
                    cnt = torch.full((1,), ids.numel(), dtype=torch.int32, device=self.device)
                keep.extend([ids, rows, cnt])
                if ids.numel() > 0:
                    t.term_ids, t.n_term_dev, t.term_rows = L.ptr(ids), L.ptr(cnt), L.ptr(rows)
        t.actions, t.actions_out = pair(tr.actions, self.actions[s], ad)
        t.rewards, t.rewards_out = pair(rewards, self.rewards[s], 1)
        t.dones, t.dones_out = pair(dones, self.dones[s], 1, torch.uint8)
        t.values, t.values_out = pair(tr.values, self.values[s], 1)
        if time_outs is not None:
            t.time_outs = L.ptr(src(time_outs, 1, torch.uint8))
            if not t.values:
                t.values = L.ptr(src(tr.values, 1))
        t.log_prob, t.log_prob_out = pair(tr.actions_log_prob, self.actions_log_prob[s], 1)
        t.mu, t.mu_out = pair(tr.action_mean, self.mu[s], ad)
        t.sigma, t.sigma_out = pair(tr.action_sigma, self.sigma[s], ad)
        L.check(L.lib.hl_record_transition(L.ctypes.byref(t), n, L.stream()))
        self.step += 1

    def add_transitions(self, transition):
        """him_rollout_storage.py:92-108: the ten slot copies of one step, as ONE launch."""
        self._record(transition, transition.rewards, transition.dones, None, transition.next_critic_observations,
                     None, None, 0.0)

    def record_env_step(self, transition, rewards, dones, infos, privileged_obs, termination_ids,
                        termination_privileged_obs, gamma, termination_count=None, assume_sorted=True):
        """The runner's terminal-observation patch (him_on_policy_runner.py:122-123),
        HIMPPO.process_env_step's time-out bootstrap (him_ppo.py:104-115) and add_transitions, fused:
        `privileged_obs` is the env's post-step privileged observation, un-patched; nothing is
        cloned.  `transition` carries what HIMPPO.act stored (observations, critic_observations,
        actions, values, actions_log_prob, action_mean, action_sigma).  `termination_ids` must be
        ascending (reset_buf.nonzero() is) unless assume_sorted=False; with `termination_count` (a
        device int32 scalar) the id / row buffers may be over-allocated and no host sync happens."""
        time_outs = infos.get("time_outs") if infos is not None else None
        self._record(transition, rewards, dones, time_outs, privileged_obs, termination_ids, termination_privileged_obs,
                     gamma, termination_count, assume_sorted)

    def clear(self):
        if self.env_writes_slots and self.step == self.num_transitions_per_env:
            # the observation the last env step produced opens the next rollout
            self._obs_all[0].copy_(self._obs_all[self.step])
            if self.privileged_observations is not None:
                self._priv_all[0].copy_(self._priv_all[self.step])
        self.step = 0

    def obs_slot(self, i):
        """(N,obs) slot i in [0, T] (env_writes_slots=True only): where the env reads (i = step) and
        writes (i = step+1) its observation history."""
        return self._obs_all[i]

    def priv_slot(self, i):
        return self._priv_all[i]

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

    def _batch_fields(self):
        """The ten (T*N, width) views in the reference's yield order (him_rollout_storage.py:176-177)."""
        flat = lambda x: x.flatten(0, 1)
        obs = flat(self.observations)
        critic = flat(self.privileged_observations) if self.privileged_observations is not None else obs
        next_critic = flat(self.next_privileged_observations) if self.next_privileged_observations is not None else obs
        return [obs, critic, flat(self.actions), next_critic, flat(self.values), flat(self.advantages), flat(self.returns),
                flat(self.actions_log_prob), flat(self.mu), flat(self.sigma)]

    def gather_batch(self, batch_idx, fields=None):
        """`x.flatten(0,1)[batch_idx]` of every rollout field as ONE launch (hl_minibatch_gather):
        the index list is read once, each row of each field is copied once."""
        fields = self._batch_fields() if fields is None else fields
        idx = batch_idx.detach().to(self.device, torch.int64).contiguous()
        m = idx.numel()
        outs = [torch.empty((m,) + tuple(x.shape[1:]), device=self.device, dtype=torch.float32) for x in fields]
        if m == 0:
            return tuple(outs)
        g = L.HlGatherFields()
        g.struct_bytes = L.ctypes.sizeof(L.HlGatherFields)
        g.n_fields = len(fields)
        srcs = []
        for k, (x, o) in enumerate(zip(fields, outs)):
            x = x if (x.is_contiguous() and x.dtype == torch.float32) else x.contiguous().float()
            srcs.append(x)
            g.src[k], g.dst[k], g.width[k] = L.ptr(x), L.ptr(o), int(x[0].numel())
        L.check(L.lib.hl_minibatch_gather(L.ctypes.byref(g), L.ptr(idx), m, int(fields[0].shape[0]), L.stream()))
        return tuple(outs)

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        """him_rollout_storage.py:137-177: same permutation draw (torch.randperm on the device),
        same slicing and yield order; the ten index gathers of a minibatch are one fused launch.
        `indices` (optional) replaces the randperm draw (tests / replay)."""
        batch = self.num_envs * self.num_transitions_per_env
        mb = batch // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        fields = self._batch_fields()
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                yield self.gather_batch(indices[i * mb:(i + 1) * mb], fields)

