"""FusedLeggedRobot: the post-physics half of `LeggedRobot`
(reference: legged_gym/legged_gym/envs/base/legged_robot.py, cited as LR) behind the same method
names, attribute names, dtypes and in-place side effects, executed by libhimloco_b200's CUDA
kernels.

What stays torch (out of scope): everything that draws torch RNG or talks to
PhysX -- `_resample_commands`, `_reset_dofs`, `_reset_root_states`, pushes, disturbances,
curricula.  They plug in through the same method names (subclass / mix in the reference class,
see INTEGRATION.md) and through `physics_step_fn` for replayed or synthetic PhysX state.

Fused step (post_physics_step):
    [torch] command resample for envs hitting the 500-step mark        LR:612-613
    hl_post_physics_fused      counters, frame, contacts, heading, 187+63-point scans,
                               termination, rewards, speculative obs + last_* roll   LR:193-241
    hl_select_reset_ids        env_ids = reset_buf.nonzero().flatten()              LR:225
    hl_terminal_rows           termination_privileged_obs, terminal_amp_states      LR:227-228
    [torch] reset_idx(env_ids)                                                      LR:229
    hl_post_reset_fixup        re-scan + obs slot 0 + roll for the reset envs       LR:232-241,332-333
"""
import ctypes
import os
from typing import Callable, Dict, Optional

import torch

from . import _lib as L
from .config import HotPathCfg


class FusedLeggedRobot:
    """Standalone env shard over replayed/synthetic PhysX tensors.  All tensors live on `device`
    (a CUDA device) and carry the reference's attribute names."""

    def __init__(self, cfg: HotPathCfg, state: Dict[str, torch.Tensor], height_samples: torch.Tensor,
                 device="cuda:0", physics_step_fn: Optional[Callable] = None, seed: int = 0):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FusedLeggedRobot runs on CUDA only (no CPU fallback)")
        n = state["root_states"].shape[0]
        self.num_envs = n
        self.num_actions = self.num_dof = self.num_dofs = 12
        self.num_bodies = cfg.num_bodies
        self.num_one_step_obs = 45
        self.num_height_points = len(cfg.measured_points_x) * len(cfg.measured_points_y)
        self.num_one_step_privileged_obs = 51 + (self.num_height_points if cfg.measure_heights else 0)
        self.num_obs, self.num_privileged_obs = 270, self.num_one_step_privileged_obs
        self.dt = cfg.dt
        self.max_episode_length = cfg.max_episode_length
        self.common_step_counter = 0
        self.extras = {}
        self.physics_step_fn = physics_step_fn
        dev = self.device
        f32 = lambda k: state[k].to(dev, torch.float32).contiguous()
        # --- PhysX-owned tensors, Isaac Gym layouts (LR:929-944)
        self.root_states = f32("root_states")
        self.dof_state = f32("dof_state")
        self.dof_pos = self.dof_state.view(n, 12, 2)[..., 0]
        self.dof_vel = self.dof_state.view(n, 12, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.rigid_body_states = f32("rigid_body_states")
        self.contact_forces = f32("contact_forces").view(n, -1, 3)
        # --- policy-side buffers (LR:954-1032)
        for k in ("actions", "last_actions", "last_last_actions", "last_dof_pos", "last_dof_vel", "torques",
                  "last_torques", "last_root_vel", "commands", "feet_air_time", "motor_strength",
                  "Kp_factors", "Kd_factors", "disturbance", "obs_buf", "privileged_obs_buf"):
            setattr(self, k, f32(k))
        self.episode_length_buf = state["episode_length_buf"].to(dev, torch.long).contiguous()
        self.terrain_levels = state["terrain_levels"].to(dev, torch.long).contiguous()
        self.last_contacts = state["last_contacts"].to(dev, torch.bool).contiguous()
        self.base_lin_vel = f32("base_lin_vel")
        self.base_ang_vel = f32("base_ang_vel")
        self.projected_gravity = f32("projected_gravity")
        self.rew_buf = torch.zeros(n, device=dev)
        self.reset_buf = torch.ones(n, dtype=torch.bool, device=dev)
        self.time_out_buf = torch.zeros(n, dtype=torch.bool, device=dev)
        self.height_samples = height_samples.to(dev, torch.int16).contiguous()
        self._initial_episode_sums = state.get("episode_sums")
        self._hl_attach(cfg, seed=seed)

    def _hl_attach(self, cfg: HotPathCfg, seed: int = 0):
        """Everything the kernels need on top of the reference's own buffers (LR:913-1032).  Also the
        hook for mixing this class into the reference's LeggedRobot: call it at
        the end of `_init_buffers` once the PhysX tensors are wrapped."""
        self.cfg_hot = cfg
        if not hasattr(self, "cfg") or self.cfg is None:
            self.cfg = cfg
        dev = self.device = torch.device(self.device)
        n = self.num_envs
        self.num_height_points = len(cfg.measured_points_x) * len(cfg.measured_points_y)
        self.num_one_step_privileged_obs = 51 + (self.num_height_points if cfg.measure_heights else 0)
        for name, shape, dt in (("contact_filt", (n, 4), torch.bool), ("feet_pos", (n, 4, 3), torch.float32),
                                ("feet_vel", (n, 4, 3), torch.float32),
                                ("measured_heights", (n, self.num_height_points), torch.float32),
                                ("joint_pos_target", (n, 12), torch.float32),
                                ("delayed_actions", (n, cfg.decimation, 12), torch.float32)):
            cur = getattr(self, name, None)
            if not (isinstance(cur, torch.Tensor) and tuple(cur.shape) == shape and cur.dtype == dt
                    and cur.is_contiguous() and cur.device == dev):
                setattr(self, name, torch.zeros(*shape, dtype=dt, device=dev))
        if self.reset_buf.dtype != torch.bool:                 # base_task.py:72 starts it as int64 ones
            self.reset_buf = self.reset_buf.to(torch.bool)
        self._base_heights = torch.zeros(n, device=dev)
        # episode sums: one (R,N) buffer; the dict holds row views so reference code keeps working
        names = cfg.episode_sum_names()
        self._episode_sums_buf = torch.zeros(max(len(names), 1), n, device=dev)
        init = getattr(self, "_initial_episode_sums", None)
        if init is not None and len(names):
            self._episode_sums_buf[:len(names)].copy_(init[:len(names)].to(dev))
        self.episode_sums = {nm: self._episode_sums_buf[k] for k, nm in enumerate(names)}
        self.reward_names, scales = cfg.active_terms()
        self.reward_scales = dict(zip(self.reward_names, scales))
        if cfg.termination_scale is not None:
            self.reward_scales["termination"] = cfg.termination_scale
        # constants the reference keeps as tensors
        t = cfg.dof_tables()
        tt = lambda a: torch.tensor(a, device=dev)
        self.default_dof_pos = tt(t["default_dof_pos"]).view(1, 12)
        self.p_gains, self.d_gains = tt(t["p_gains"]), tt(t["d_gains"])
        self.torque_limits, self.dof_vel_limits = tt(t["torque_limits"]), tt(t["dof_vel_limits"])
        self.dof_pos_limits = torch.stack([tt(t["dof_pos_lo"]), tt(t["dof_pos_hi"])], dim=-1)
        self.noise_scale_vec = tt(cfg.noise_scale_vec())
        self.add_noise = cfg.add_noise
        self.feet_indices = torch.tensor(cfg.feet_indices, dtype=torch.long, device=dev)
        self.penalised_contact_indices = torch.tensor(cfg.penalised_contact_indices, dtype=torch.long, device=dev)
        self.termination_contact_indices = torch.tensor(cfg.termination_contact_indices, dtype=torch.long, device=dev)
        # terrain: the one-off min-of-3 table
        self._height_min3 = None
        self._height_min3f = None
        # reset-id compaction + terminal rows (capacity N; counts live on the device)
        self._reset_ids = torch.zeros(n, dtype=torch.long, device=dev)
        self._n_reset = torch.zeros(1, dtype=torch.int32, device=dev)
        self._term_priv = torch.zeros(n, self.num_one_step_privileged_obs, device=dev)
        self._term_amp = torch.zeros(n, 30, device=dev)
        self._select_ws = torch.zeros(int(L.lib.hl_select_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        self._fused_ws = torch.zeros(int(L.lib.hl_fused_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        self._selterm_ws = torch.zeros(int(L.lib.hl_select_terminal_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        # ids + terminal rows straight from the fused kernel: wins in the launch-bound small-N regime,
        # costs more than the separate compaction launch at large N (measured)
        self.single_launch = n <= 16384
        if os.environ.get("HL_SINGLE_LAUNCH") in ("0", "1"):      # A/B knob
            self.single_launch = os.environ["HL_SINGLE_LAUNCH"] == "1"
        # noise: Philox by default; parity tests install pre-drawn tensors
        self._noise = {}
        self._philox_seed = int(seed)
        self._c = cfg.to_c()
        self._bufs = None
        self._fused_event_hook = None      # bench.py: CUDA-event pair around the fused kernel
        self._rollout = None               # bind_rollout(): storage whose slots the env writes in place
        self._prepare_terrain()

    # ------------------------------------------------------------------ plumbing
    def _prepare_terrain(self):
        if self.cfg_hot.is_plane:
            return
        rows, cols = self.height_samples.shape
        if (rows, cols) != tuple(self.cfg_hot.terrain_shape):
            raise ValueError(f"height_samples is {rows}x{cols}, cfg expects {self.cfg_hot.terrain_shape}")
        self._height_min3 = torch.empty(rows - 1, cols - 1, dtype=torch.int16, device=self.device)
        L.check(L.lib.hl_terrain_prepare(L.ptr(self.height_samples), rows, cols, L.ptr(self._height_min3), L.stream()))
        # the same table in metres (fp32): one gather per scan point in the persistent fused kernel
        self._height_min3f = torch.empty(rows - 1, cols - 1, dtype=torch.float32, device=self.device)
        L.check(L.lib.hl_terrain_prepare_f32(L.ptr(self.height_samples), rows, cols, float(self.cfg_hot.vertical_scale),
                                             L.ptr(self._height_min3f), L.stream()))

    def set_noise_tensors(self, obs45=None, obs187=None, term45=None, term187=None):
        """Parity mode: the U[0,1) draws `torch.rand_like` would return at LR:394,400 (obs) and
        LR:451,457 (terminal obs).  None => in-kernel Philox."""
        cv = lambda t: None if t is None else t.to(self.device, torch.float32).contiguous()
        self._noise = dict(obs45=cv(obs45), obs187=cv(obs187), term45=cv(term45), term187=cv(term187))
        self._bufs = None

    def refresh_buffers(self):
        """Call after rebinding any tensor attribute (pointers are cached in a C struct)."""
        self._bufs = None

    # ------------------------------------------------------------------ zero-copy rollout slots
    def bind_rollout(self, storage):
        """Make the env read its observation history from rollout slot `storage.step` and write the
        new observations / privileged observations into slot step+1 of a
        HIMRolloutStorage(env_writes_slots=True): add_transitions' two largest copies
        (him_rollout_storage.py:97-98) disappear because transition.observations and
        .critic_observations already ARE the slots.  bind_rollout(None) detaches."""
        if storage is None:
            if getattr(self, "_rollout", None) is not None:
                self.obs_buf, self.privileged_obs_buf = self.obs_buf.clone(), self.privileged_obs_buf.clone()
            self._rollout = None
            self._bufs = None
            return
        if not getattr(storage, "env_writes_slots", False):
            raise ValueError("bind_rollout needs HIMRolloutStorage(env_writes_slots=True)")
        s = storage.step
        storage.obs_slot(s).copy_(self.obs_buf)
        storage.priv_slot(s).copy_(self.privileged_obs_buf)
        self.obs_buf, self.privileged_obs_buf = storage.obs_slot(s), storage.priv_slot(s)
        self._rollout = storage
        self._bufs = None

    def _bind_rollout_slot(self, bufs):
        st = self._rollout
        s = st.step
        if s >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")     # as add_transitions would (him_rollout_storage.py:94)
        self.obs_buf, self.privileged_obs_buf = st.obs_slot(s + 1), st.priv_slot(s + 1)
        bufs.obs_buf_in = L.ptr(st.obs_slot(s))
        bufs.obs_buf_out = L.ptr(self.obs_buf)
        bufs.privileged_obs_buf = L.ptr(self.privileged_obs_buf)

    def _buffers(self) -> L.HlEnvBuffers:
        if self._bufs is not None:
            self._bufs.philox_offset = self.common_step_counter
            return self._bufs
        b = L.HlEnvBuffers()
        b.struct_bytes = ctypes.sizeof(L.HlEnvBuffers)
        p = L.ptr
        b.root_states, b.dof_state = p(self.root_states), p(self.dof_state)
        b.contact_forces, b.rigid_body_states = p(self.contact_forces), p(self.rigid_body_states)
        b.height_samples = p(self.height_samples)
        b.height_min3 = p(self._height_min3)
        b.height_min3f = p(self._height_min3f)
        b.fused_ws = p(self._fused_ws)        # tile ticket of the persistent kernel (+ look-back states)
        b.actions, b.last_actions, b.last_last_actions = p(self.actions), p(self.last_actions), p(self.last_last_actions)
        b.last_dof_pos, b.last_dof_vel = p(self.last_dof_pos), p(self.last_dof_vel)
        b.torques, b.last_torques, b.last_root_vel = p(self.torques), p(self.last_torques), p(self.last_root_vel)
        b.commands, b.episode_length_buf = p(self.commands), p(self.episode_length_buf)
        b.last_contacts, b.contact_filt = p(self.last_contacts), p(self.contact_filt)
        b.feet_air_time, b.disturbance = p(self.feet_air_time), p(self.disturbance)
        b.terrain_levels, b.episode_sums = p(self.terrain_levels), p(self._episode_sums_buf)
        b.base_lin_vel, b.base_ang_vel = p(self.base_lin_vel), p(self.base_ang_vel)
        b.projected_gravity, b.measured_heights = p(self.projected_gravity), p(self.measured_heights)
        b.feet_pos, b.feet_vel = p(self.feet_pos), p(self.feet_vel)
        b.reset_buf, b.time_out_buf, b.rew_buf = p(self.reset_buf), p(self.time_out_buf), p(self.rew_buf)
        b.obs_buf_in = b.obs_buf_out = p(self.obs_buf)
        b.privileged_obs_buf = p(self.privileged_obs_buf)
        b.noise_u45, b.noise_u187 = p(self._noise.get("obs45")), p(self._noise.get("obs187"))
        b.philox_seed, b.philox_offset = self._philox_seed, self.common_step_counter
        b.height_idx_out = None
        b.base_height_out = p(self._base_heights)
        if self.single_launch:
            b.reset_ids_out, b.n_reset_out = p(self._reset_ids), p(self._n_reset)
            b.term_priv_out, b.term_amp_out = p(self._term_priv), p(self._term_amp)
            b.term_noise_u45, b.term_noise_u187 = p(self._noise.get("term45")), p(self._noise.get("term187"))
        self._bufs = b
        return b

    def _stages(self, stages: int, env_ids: Optional[torch.Tensor] = None, n_ids: Optional[torch.Tensor] = None):
        if env_ids is not None and n_ids is None:
            if env_ids.numel() == 0:
                return
            env_ids = env_ids.to(self.device, torch.long).contiguous()
            n_ids = torch.tensor([env_ids.numel()], dtype=torch.int32, device=self.device)
        L.check(L.lib.hl_post_physics_stages(ctypes.byref(self._c), ctypes.byref(self._buffers()), stages,
                                             L.ptr(env_ids), L.ptr(n_ids), self.num_envs, L.stream()))

    # ------------------------------------------------------------------ the reference's methods
    def _compute_torques(self, actions):
        """LR:658-688.  `actions` may be a column slice of delayed_actions (row-strided)."""
        if actions.stride(-1) != 1:
            actions = actions.contiguous()
        out = torch.empty(self.num_envs, 12, device=self.device)
        L.check(L.lib.hl_pd_torque(ctypes.byref(self._c), L.ptr(actions), actions.stride(0), L.ptr(self.dof_state),
                                   L.ptr(self.motor_strength), L.ptr(self.Kp_factors), L.ptr(self.Kd_factors),
                                   L.ptr(self.last_dof_vel), L.ptr(out), L.ptr(self.joint_pos_target), self.num_envs,
                                   L.stream()))
        return out

    def _compute_torques_into(self, actions, out):
        """Same, writing into a preallocated (N,12) tensor (no allocation; graph-capturable)."""
        L.check(L.lib.hl_pd_torque(ctypes.byref(self._c), L.ptr(actions), actions.stride(0), L.ptr(self.dof_state),
                                   L.ptr(self.motor_strength), L.ptr(self.Kp_factors), L.ptr(self.Kd_factors),
                                   L.ptr(self.last_dof_vel), L.ptr(out), L.ptr(self.joint_pos_target), self.num_envs,
                                   L.stream()))
        return out

    def _get_heights(self, env_ids=None):
        """LR:1318-1355 -> (N,187); the scan itself always covers every env like the reference's
        `env_ids=None` path (the only one it ever takes)."""
        self._stages(L.ST_HEIGHTS)
        return self.measured_heights

    def _get_base_heights(self, env_ids=None):
        """LR:1357-1398 -> (N,)."""
        self._stages(L.ST_BASE_HEIGHT)
        return self._base_heights

    def get_height_indices(self):
        """Debug/parity: the clipped (px,py) cell of every scan point, (N,187,2) int32."""
        idx = torch.empty(self.num_envs, self.num_height_points, 2, dtype=torch.int32, device=self.device)
        b = self._buffers()
        b.height_idx_out = L.ptr(idx)
        try:
            self._stages(L.ST_HEIGHTS)
        finally:
            b.height_idx_out = None
        return idx

    def update_base_frame(self):
        """LR:197-209: base_lin_vel, base_ang_vel, projected_gravity, feet_pos/vel, contacts."""
        self._stages(L.ST_FRAME | L.ST_CONTACTS)

    def update_heading_command(self):
        """LR:616-620."""
        self._stages(L.ST_HEADING)

    def check_termination(self):
        """LR:249-286 (sets reset_buf, time_out_buf; the dead .item() bookkeeping is dropped)."""
        self._stages(L.ST_TERMINATION)

    def compute_reward(self):
        """LR:363-380."""
        self._stages(L.ST_REWARD)

    def compute_observations(self):
        """LR:382-404."""
        self._stages(L.ST_OBS)

    def compute_termination_observations(self, env_ids):
        """LR:439-460 -> (len(env_ids), 238)."""
        env_ids = env_ids.to(self.device, torch.long).contiguous()
        if env_ids.numel() == 0:
            return self._term_priv[:0]
        n_ids = torch.tensor([env_ids.numel()], dtype=torch.int32, device=self.device)
        self._terminal_rows(env_ids, n_ids)
        return self._term_priv[:env_ids.numel()]

    def get_amp_observations(self):
        """LR:406-416 -> (N,30)."""
        out = torch.empty(self.num_envs, 30, device=self.device)
        L.check(L.lib.hl_amp_observations(L.ptr(self.dof_state), L.ptr(self.base_lin_vel), L.ptr(self.base_ang_vel),
                                          L.ptr(out), self.num_envs, L.stream()))
        return out

    def _terminal_rows(self, env_ids, n_ids):
        L.check(L.lib.hl_terminal_rows(ctypes.byref(self._c), ctypes.byref(self._buffers()), L.ptr(env_ids), L.ptr(n_ids),
                                       L.ptr(self._noise.get("term45")), L.ptr(self._noise.get("term187")),
                                       L.ptr(self._term_priv), L.ptr(self._term_amp), self.num_envs, L.stream()))

    # ------------------------------------------------------------------ fused step pieces
    def fused_pre_reset(self):
        """Launch the fused kernel, the id compaction and the terminal rows; no host sync."""
        bufs = self._buffers()
        if getattr(self, "_rollout", None) is not None:
            self._bind_rollout_slot(bufs)
        c, b = ctypes.byref(self._c), ctypes.byref(bufs)
        ev = self._fused_event_hook() if self._fused_event_hook is not None else None
        L.check(L.lib.hl_post_physics_fused(c, b, self.num_envs, L.stream()))
        if ev is not None:
            ev()
        bufs.flags |= 1          # HL_BUF_HISTORY_CLIPPED: the step just clipped the whole obs_buf (LR:168)
        if self.single_launch:
            return               # ids, count and terminal rows came out of the same launch
        L.check(L.lib.hl_select_and_terminal(c, b, L.ptr(self._noise.get("term45")), L.ptr(self._noise.get("term187")),
                                             L.ptr(self._reset_ids), L.ptr(self._n_reset), L.ptr(self._term_priv),
                                             L.ptr(self._term_amp), L.ptr(self._selterm_ws), self.num_envs, L.stream()))

    def fused_post_reset(self, with_reset_zero: bool = False):
        """Patch the reset envs after reset_idx.  with_reset_zero=True also applies reset_idx's own
        RNG-free buffer resets in the kernel (for replay loops that skip the torch reset_idx)."""
        L.check(L.lib.hl_post_reset_fixup(ctypes.byref(self._c), ctypes.byref(self._buffers()), L.ptr(self._reset_ids),
                                          L.ptr(self._n_reset), int(with_reset_zero), self.num_envs, L.stream()))

    def post_physics_step(self):
        """LR:178-247 -> (env_ids, termination_privileged_obs, terminal_amp_states)."""
        self._pre_step_callbacks()
        self.fused_pre_reset()
        n_reset = int(self._n_reset.item())          # the reference syncs here too (LR:225,298)
        env_ids = self._reset_ids[:n_reset]
        term_priv, term_amp = self._term_priv[:n_reset], self._term_amp[:n_reset]
        self.reset_idx(env_ids)
        self.fused_post_reset()
        self.common_step_counter += 1
        return env_ids, term_priv, term_amp

    def step(self, actions):
        """LR:122-176 (7-tuple; the AMP runner reads terminal_amp_states off `extras`)."""
        clip = self.cfg_hot.clip_actions
        torch.clamp(actions.to(self.device), -clip, clip, out=self.actions)
        self._delay_actions()
        for k in range(self.cfg_hot.decimation):
            self._compute_torques_into(self.delayed_actions[:, k], self.torques)
            if self.physics_step_fn is not None:
                self.physics_step_fn(self, k)
        env_ids, term_priv, term_amp = self.post_physics_step()
        self.extras["terminal_amp_states"] = term_amp
        return (self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras, env_ids, term_priv)

    # ------------------------------------------------------------------ torch-side hooks (RNG / PhysX; out of scope)
    def _delay_actions(self):
        """LR:133-138 (torch RNG draw of the per-env action delay)."""
        dec = self.cfg_hot.decimation
        delay = torch.randint(0, dec, (self.num_envs, 1), device=self.device)
        steps = torch.arange(dec, device=self.device).view(1, dec, 1)
        mask = (steps >= delay.view(-1, 1, 1)).to(self.actions.dtype)
        torch.add(self.last_actions.unsqueeze(1), (self.actions - self.last_actions).unsqueeze(1) * mask,
                  out=self.delayed_actions)

    def _pre_step_callbacks(self):
        """The RNG-driven part of _post_physics_step_callback (LR:612-613): resample commands of
        envs whose *incremented* episode length hits the interval.  Hook; default = keep."""
        return None

    def _reset_dofs(self, env_ids):
        return None

    def _reset_root_states(self, env_ids):
        return None

    def _resample_commands(self, env_ids):
        return None

    def reset_idx(self, env_ids):
        """The deterministic bookkeeping of LR:288-361; state re-draws go through the three hooks
        above.  The full-N height re-scan of LR:332-333 is replaced by the targeted re-scan in
        hl_post_reset_fixup (identical result: non-reset envs did not move)."""
        if len(env_ids) == 0:
            return
        self._reset_dofs(env_ids)
        self._reset_root_states(env_ids)
        self._resample_commands(env_ids)
        self.last_actions[env_ids] = 0.0
        self.last_last_actions[env_ids] = 0.0
        self.last_dof_pos[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.last_torques[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.reset_buf[env_ids] = True
        self.extras["episode"] = {}
        if self.episode_sums:   # LR:346-350 as one launch (means of every row over the reset ids, rows zeroed)
            rows = self._episode_sums_buf.shape[0]
            means = torch.empty(rows, device=self.device)
            ids = env_ids.to(self.device, torch.int64).contiguous()
            cnt = torch.full((1,), ids.numel(), dtype=torch.int32, device=self.device)
            L.check(L.lib.hl_episode_means(L.ptr(self._episode_sums_buf), L.ptr(self.episode_length_buf), L.ptr(ids),
                                           L.ptr(cnt), rows, self.num_envs, float(self.dt), 1, L.ptr(means), L.stream()))
            for k, key in enumerate(self.episode_sums.keys()):
                self.extras["episode"]["rew_" + key] = means[k]
        self.extras["time_outs"] = self.time_out_buf
        self.episode_length_buf[env_ids] = 0

    def snapshot(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k in ("base_lin_vel", "base_ang_vel", "projected_gravity", "measured_heights", "reset_buf",
                  "time_out_buf", "rew_buf", "contact_filt", "last_contacts", "feet_air_time", "commands",
                  "episode_length_buf", "obs_buf", "privileged_obs_buf", "last_actions", "last_last_actions",
                  "last_dof_pos", "last_dof_vel", "last_torques", "last_root_vel"):
            out[k] = getattr(self, k).detach().cpu().clone()
        if self.episode_sums:
            out["episode_sums"] = self._episode_sums_buf[:len(self.episode_sums)].detach().cpu().clone()
        return out
