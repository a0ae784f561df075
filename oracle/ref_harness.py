"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Tier-0 oracle: runs the *actual* reference classes from /root/reference on CPU under
stubs for the third-party modules that are absent from this image (isaacgym Preview 4,
pybullet_utils, matplotlib, ruamel.yaml).  Works only in the build container (the GPU box has
no /root/reference); it is used by oracle/make_goldens.py to mint tests/golden/*.npz and by
tests/test_oracle_vs_reference.py (skipped when the reference is absent) to validate the
Tier-1 restatement in oracle/torch_oracle.py.

The stubbed `isaacgym.torch_utils` helpers restate the public Isaac Gym definitions
(SURVEY.md Appendix B.2).  isaacgym is a proprietary wheel that is not vendored by the
reference and not installable offline => parity is unpinned at that boundary: nothing in the
reference repo pins those 7 helpers.
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("HIMLOCO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "legged_gym", "legged_gym"))


# --------------------------------------------------------------------------- stubs
def _torch_utils_stub():
    m = types.ModuleType("isaacgym.torch_utils")

    def quat_rotate_inverse(q, v):
        shape = q.shape
        q_w = q[:, -1]
        q_vec = q[:, :3]
        a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
        b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
        c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
        return a - b + c

    def quat_apply(a, b):
        shape = b.shape
        a = a.reshape(-1, 4)
        b = b.reshape(-1, 3)
        xyz = a[:, :3]
        t = xyz.cross(b, dim=-1) * 2
        return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)

    def normalize(x, eps: float = 1e-9):
        return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)

    def torch_rand_float(lower, upper, shape, device):
        return (upper - lower) * torch.rand(*shape, device=device) + lower

    def to_torch(x, dtype=torch.float, device="cuda:0", requires_grad=False):
        return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)

    def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
        zs = np.zeros((n_dims,))
        assert axis_idx < n_dims
        zs[axis_idx] = 1.0
        params = np.where(zs == 1.0, value, zs)
        params[0] = x_value
        return list(params.astype(dtype))

    def quat_from_euler_xyz(roll, pitch, yaw):
        cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
        cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
        cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
        qw = cy * cr * cp + sy * sr * sp
        qx = cy * sr * cp - sy * cr * sp
        qy = cy * cr * sp + sy * sr * cp
        qz = sy * cr * cp - cy * sr * sp
        return torch.stack([qx, qy, qz, qw], dim=-1)

    for f in (quat_rotate_inverse, quat_apply, normalize, torch_rand_float, to_torch,
              get_axis_params, quat_from_euler_xyz):
        setattr(m, f.__name__, f)
    # `from isaacgym.torch_utils import *` also has to provide these names
    # (the real module also leaks `np` and `torch` through the star import; legged_robot.py:1261
    # relies on that for `np.ceil`)
    m.np, m.torch = np, torch
    m.__all__ = ["quat_rotate_inverse", "quat_apply", "normalize", "torch_rand_float", "to_torch",
                 "get_axis_params", "quat_from_euler_xyz", "np", "torch"]
    return m


class _Anything:
    """Attribute sink: any attribute is another sink; calling returns None."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return None


def install_stubs():
    if "isaacgym" in sys.modules and getattr(sys.modules["isaacgym"], "_himloco_stub", False):
        return
    if not hasattr(np, "int"):
        np.int = int  # numpy 2.x removed it; used at motion_loader.py:210,234
    if not hasattr(np, "float"):
        np.float = float
    ig = types.ModuleType("isaacgym")
    ig._himloco_stub = True
    ig.torch_utils = _torch_utils_stub()
    gymtorch = types.ModuleType("isaacgym.gymtorch")
    gymtorch.unwrap_tensor = lambda t: t
    gymtorch.wrap_tensor = lambda t: t
    gymapi = types.ModuleType("isaacgym.gymapi")
    gymapi.__getattr__ = lambda name: _Anything()
    gymutil = types.ModuleType("isaacgym.gymutil")
    gymutil.__getattr__ = lambda name: _Anything()
    terrain_utils = types.ModuleType("isaacgym.terrain_utils")
    terrain_utils.__getattr__ = lambda name: _Anything()
    ig.gymtorch, ig.gymapi, ig.gymutil, ig.terrain_utils = gymtorch, gymapi, gymutil, terrain_utils
    sys.modules.update({
        "isaacgym": ig, "isaacgym.torch_utils": ig.torch_utils, "isaacgym.gymtorch": gymtorch,
        "isaacgym.gymapi": gymapi, "isaacgym.gymutil": gymutil,
        "isaacgym.terrain_utils": terrain_utils,
    })
    for name in ("matplotlib", "matplotlib.pyplot", "ruamel", "ruamel.yaml", "pybullet_utils",
                 "pybullet_utils.transformations", "tensorboard", "torch.utils.tensorboard"):
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__getattr__ = lambda n: _Anything()
            sys.modules[name] = mod
    for p in (os.path.join(REFERENCE_ROOT, "rsl_rl"), os.path.join(REFERENCE_ROOT, "legged_gym")):
        if p not in sys.path:
            sys.path.insert(0, p)


class FakeGym:
    def __getattr__(self, name):
        return lambda *a, **k: None


# --------------------------------------------------------------------------- env builder
def reference_cfg(task: str):
    """Instantiate the reference's own config class for `task` in {flat, stairs, amp, recover}."""
    install_stubs()
    import importlib
    if task == "flat":
        mod = importlib.import_module("legged_gym.envs.aliengo.aliengo_config")
        cfg = mod.AlienGoRoughCfg()
    elif task == "stairs":
        mod = importlib.import_module("legged_gym.envs.aliengo.aliengo_stairs_config")
        cfg = mod.AlienGoStairsCfg()
    elif task == "amp":
        mod = importlib.import_module("legged_gym.envs.aliengo.aliengo_amp_config")
        cfg = mod.AlienGoRoughCfg()
    elif task == "recover":
        mod = importlib.import_module("legged_gym.envs.aliengo.aliengo_recover_config")
        cfg = mod.AlienGoRoughRecoverCfg()
    else:
        raise ValueError(task)
    return cfg


def build_reference_env(task: str, state: dict, height_samples: torch.Tensor, sum_names=None, hot_cfg=None):
    """LeggedRobot.__new__ + hand-set attributes.

    `state` is the dict produced by isaacgymloco_b200.synthetic.make_state (CPU tensors);
    every tensor is cloned so the reference may mutate freely.
    """
    install_stubs()
    from legged_gym.envs.base.legged_robot import LeggedRobot

    cfg = reference_cfg(task)
    n = state["root_states"].shape[0]
    cfg.env.num_envs = n
    env = LeggedRobot.__new__(LeggedRobot)
    env.cfg = cfg
    env.device = "cpu"
    env.num_envs = n
    env.num_actions = env.num_dof = env.num_dofs = 12
    env.num_bodies = 17
    env.num_obs = cfg.env.num_observations
    env.num_privileged_obs = cfg.env.num_privileged_obs
    env.num_one_step_obs = cfg.env.num_one_step_observations
    env.num_one_step_privileged_obs = cfg.env.num_one_step_privileged_obs
    env.history_length = int(env.num_obs / env.num_one_step_obs)
    env.gym, env.sim, env.viewer = FakeGym(), None, None
    env.enable_viewer_sync, env.debug_viz, env.headless = False, False, True
    env.sim_params = types.SimpleNamespace(dt=0.005)
    env.up_axis_idx = 2
    props = cfg.terrain.terrain_proportions
    import math
    tp = list(props) + [0.0] * 10
    env.stairsup_start_idx = math.ceil(n * sum(tp[:4]))
    env.stairsup_end_idx = math.ceil(n * sum(tp[:5]))
    env.pit_start_idx = math.ceil(n * sum(tp[:8]))
    env.gap_end_idx = n
    env._parse_cfg(cfg)

    c = lambda k: state[k].clone()
    env.obs_buf = c("obs_buf")
    env.privileged_obs_buf = c("privileged_obs_buf")
    env.rew_buf = torch.zeros(n)
    env.reset_buf = torch.ones(n, dtype=torch.long)
    env.episode_length_buf = c("episode_length_buf")
    env.time_out_buf = torch.zeros(n, dtype=torch.bool)
    env.extras = {}
    env.common_step_counter = 0
    env.init_done = True

    tc = cfg.terrain
    terrain = types.SimpleNamespace(
        cfg=tc, xSize=tc.terrain_length * tc.num_rows, ySize=tc.terrain_width * tc.num_cols,
        env_length=tc.terrain_length, env_width=tc.terrain_width)
    from legged_gym.utils.terrain import Terrain
    terrain.in_terrain_range = types.MethodType(Terrain.in_terrain_range, terrain)
    env.terrain = terrain
    env.height_samples = height_samples.clone()

    env.root_states = c("root_states")
    env.dof_state = c("dof_state")
    env.dof_pos = env.dof_state.view(n, 12, 2)[..., 0]
    env.dof_vel = env.dof_state.view(n, 12, 2)[..., 1]
    env.base_quat = env.root_states[:, 3:7]
    env.rigid_body_states = c("rigid_body_states")
    env.contact_forces = c("contact_forces").view(n, -1, 3)
    # body index lists as _create_envs would build them (LR:1151-1153,1213-1215) for this cfg
    env.feet_indices = torch.tensor([4, 8, 12, 16])
    env.penalised_contact_indices = torch.tensor([2, 6, 10, 14, 3, 7, 11, 15, 0])
    env.termination_contact_indices = torch.tensor(
        [0] if len(cfg.asset.terminate_after_contacts_on) else [], dtype=torch.long)
    if hot_cfg is not None:
        assert list(env.termination_contact_indices) == list(hot_cfg.termination_contact_indices)
        assert bool(cfg.commands.heading_command) == bool(hot_cfg.heading_command)
        assert list(cfg.terrain.terrain_proportions) == list(hot_cfg.terrain_proportions)
    env.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(n, 1)
    env.forward_vec = torch.tensor([1.0, 0.0, 0.0]).repeat(n, 1)
    env.torques = c("torques")
    env.p_gains = torch.full((12,), 40.0)
    env.d_gains = torch.full((12,), 2.0)
    env.actions = c("actions")
    env.last_actions = c("last_actions")
    env.last_last_actions = c("last_last_actions")
    env.last_dof_pos = c("last_dof_pos")
    env.last_dof_vel = c("last_dof_vel")
    env.last_torques = c("last_torques")
    env.last_root_vel = c("last_root_vel")
    env.commands = c("commands")
    env.commands_scale = torch.tensor([env.obs_scales.lin_vel, env.obs_scales.lin_vel,
                                       env.obs_scales.ang_vel])
    env.feet_air_time = c("feet_air_time")
    env.last_contacts = c("last_contacts")
    env.base_lin_vel = c("base_lin_vel")
    env.base_ang_vel = c("base_ang_vel")
    env.projected_gravity = c("projected_gravity")
    env.default_dof_pos = c("default_dof_pos").view(1, 12)
    env.motor_strength = c("motor_strength")
    env.Kp_factors = c("Kp_factors")
    env.Kd_factors = c("Kd_factors")
    env.motor_strength_factors = torch.ones(n, 1)
    env.disturbance = c("disturbance")
    env.terrain_levels = c("terrain_levels")
    env.dof_pos_limits = c("dof_pos_limits")
    env.dof_vel_limits = c("dof_vel_limits")
    env.torque_limits = c("torque_limits")
    env.noise_scale_vec = env._get_noise_scale_vec(cfg)
    env.height_points = env._init_height_points()
    env.base_height_points = env._init_base_height_points()
    env.measured_heights = env._get_heights()
    env._prepare_reward_function()
    # rows of state["episode_sums"] follow `sum_names` (HotPathCfg.episode_sum_names())
    sum_names = list(sum_names) if sum_names is not None else list(env.episode_sums.keys())
    assert sorted(sum_names) == sorted(env.episode_sums.keys()), (sum_names, list(env.episode_sums))
    for k, name in enumerate(sum_names):
        env.episode_sums[name] = state["episode_sums"][k].clone()
    return env


def load_reference_amp_loader(motion_files=None, preload=False, num_preload=0, dt=0.02):
    install_stubs()
    import glob
    from rsl_rl.datasets.motion_loader import AMPLoader
    if motion_files is None:
        motion_files = default_aliengo_motion_files()
    return AMPLoader(device="cpu", time_between_frames=dt, preload_transitions=preload,
                     num_preload_transitions=num_preload, motion_files=motion_files)


def default_aliengo_motion_files():
    """The glob the AMP cfg uses (aliengo_amp_config.py:34-36): trot*, left*, right*."""
    import glob
    d = os.path.join(REFERENCE_ROOT, "datasets", "mocap_motions_aliengo")
    files = []
    for pat in ("trot*", "left*", "right*"):
        files += sorted(glob.glob(os.path.join(d, pat)))
    return files
