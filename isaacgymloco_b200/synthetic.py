"""Seeded synthetic PhysX state, terrain and rollout tensors.

PhysX itself is out of scope: these tensors stand in for what `gym.refresh_*_tensor` would
expose, in Isaac Gym's own AoS layouts (legged_robot.py:929-944):
root_states (N,13), dof_state (N*12,2), rigid_body_states (N*17,13), contact_forces (N*17,3).
Everything is generated on CPU with torch/numpy generators and moved to `device` afterwards so
the CPU oracle and the CUDA path see bit-identical inputs.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch

from .config import HotPathCfg


# ----------------------------------------------------------------------------- terrain
def make_terrain(cfg: HotPathCfg, seed: int = 0, kind: str = "auto") -> torch.Tensor:
    """Synthetic int16 height field in the reference's format (terrain.py:55-62: shape
    (tot_rows, tot_cols), x -> rows, raw units of `vertical_scale`, `border` flat cells around).

    Tiles are laid out like Terrain.curiculum (column = type by cumulative proportion, row =
    difficulty): uniform roughness, smooth/rough slopes and up/down pyramid stairs.  The generators
    are written here (they are not the reference's), only the format is the reference's.
    """
    rows, cols = cfg.terrain_shape
    hf = np.zeros((rows, cols), dtype=np.int16)
    if cfg.is_plane:
        return torch.from_numpy(hf)
    rng = np.random.default_rng(seed)
    hs, vs = cfg.horizontal_scale, cfg.vertical_scale
    border = int(cfg.border_size / hs)
    lpp = int(cfg.terrain_length / hs)
    wpp = int(cfg.terrain_width / hs)
    props = np.cumsum(list(cfg.terrain_proportions) + [0.0] * 10)[:10]
    for j in range(cfg.num_cols):
        choice = j / cfg.num_cols + 0.001
        for i in range(cfg.num_rows):
            difficulty = i / cfg.num_rows
            tile = np.zeros((lpp, wpp), dtype=np.float64)      # metres
            xs = (np.arange(lpp) - lpp / 2) * hs
            ys = (np.arange(wpp) - wpp / 2) * hs
            rad = np.maximum(np.abs(xs)[:, None], np.abs(ys)[None, :])  # pyramid distance
            if choice < props[0]:
                pass                                                   # flat
            elif choice < props[1]:
                tile += rng.uniform(-0.02 - 0.06 * difficulty, 0.02 + 0.06 * difficulty, tile.shape)
            elif choice < props[2] or choice < props[3]:
                slope = 0.4 * difficulty * (1 if j % 2 else -1)
                plat = 1.0
                tile += slope * np.maximum(0.0, (min(lpp, wpp) * hs / 2 - plat) - np.maximum(rad - plat, 0))
                if choice >= props[2]:
                    tile += rng.uniform(-0.03, 0.03, tile.shape)
            elif choice < props[5]:
                step_h = (0.05 + 0.18 * difficulty) * (1 if choice < props[4] else -1)
                step_w = 0.30
                n_steps = np.floor(np.maximum(0.0, (min(lpp, wpp) * hs / 2 - 1.5) - np.maximum(rad - 1.5, 0)) / step_w)
                tile += n_steps * step_h
            else:
                amp = 0.05 + 0.15 * difficulty
                blocks = rng.uniform(-amp, amp, (lpp // 8 + 1, wpp // 8 + 1))
                tile += np.kron(blocks, np.ones((8, 8)))[:lpp, :wpp]
            x0, y0 = border + i * lpp, border + j * wpp
            hf[x0:x0 + lpp, y0:y0 + wpp] = np.round(tile / vs).astype(np.int16)
    return torch.from_numpy(hf)


def _terrain_height_at(cfg: HotPathCfg, hf: torch.Tensor, xy: torch.Tensor) -> torch.Tensor:
    if cfg.is_plane:
        return torch.zeros(xy.shape[0])
    ix = ((xy[:, 0] + cfg.border_size) / cfg.horizontal_scale).long().clamp(0, hf.shape[0] - 1)
    iy = ((xy[:, 1] + cfg.border_size) / cfg.horizontal_scale).long().clamp(0, hf.shape[1] - 1)
    return hf[ix, iy].float() * cfg.vertical_scale


def _quat_from_euler_xyz(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                        sy * cr * cp - cy * sr * sp, cy * cr * cp + sy * sr * sp], dim=-1)


# ----------------------------------------------------------------------------- env state
def make_state(cfg: HotPathCfg, n: int, hf: torch.Tensor, seed: int = 1234, step: int = 0,
               env_id_offset: int = 0) -> Dict[str, torch.Tensor]:
    """One synthetic post-PhysX snapshot for `n` envs (CPU tensors, reference attribute names)."""
    g = torch.Generator().manual_seed(seed + 7919 * step + 104729 * env_id_offset)
    randn = lambda *s: torch.randn(*s, generator=g)
    rand = lambda *s: torch.rand(*s, generator=g)
    tabs = cfg.dof_tables()
    default = torch.from_numpy(tabs["default_dof_pos"])
    s: Dict[str, torch.Tensor] = {}

    # --- root
    xs, ys = (cfg.x_size, cfg.y_size) if not cfg.is_plane else (80.0, 160.0)
    xy = torch.stack([rand(n) * xs, rand(n) * ys], dim=-1)
    outside = rand(n) < 0.02
    shift = torch.where(rand(n) < 0.5, torch.tensor(-1.0), torch.tensor(1.0))
    xy[:, 0] = torch.where(outside & (shift < 0), -rand(n) * 5.0, xy[:, 0])
    xy[:, 1] = torch.where(outside & (shift > 0), ys + cfg.border_size / 2 + rand(n) * 5.0, xy[:, 1])
    z = _terrain_height_at(cfg, hf, xy) + 0.42 + 0.03 * randn(n)
    quat = _quat_from_euler_xyz(0.1 * randn(n), 0.1 * randn(n), (rand(n) * 2 - 1) * math.pi)
    lin = 0.5 * randn(n, 3)
    lin[:, 2] = torch.where(rand(n) < 0.001, -5.0 - rand(n), lin[:, 2])
    ang = 0.5 * randn(n, 3)
    s["root_states"] = torch.cat([xy, z[:, None], quat, lin, ang], dim=-1).contiguous()

    # --- dofs (interleaved pos, vel)
    dof_pos = default[None, :] + 0.2 * randn(n, 12)
    dof_vel = 2.0 * randn(n, 12)
    s["dof_state"] = torch.stack([dof_pos, dof_vel], dim=-1).reshape(n * 12, 2).contiguous()

    # --- contacts
    nb = cfg.num_bodies
    cf = torch.zeros(n, nb, 3)
    feet = cfg.feet_indices
    in_contact = rand(n, 4) < 0.5
    fz = (60 + 40 * randn(n, 4)).abs() * in_contact
    fxy = 5.0 * randn(n, 4, 2) * in_contact[..., None]
    # a few stumbling feet (|Fxy| > 5 |Fz|) so feet_stumble is exercised
    stumble = (rand(n, 4) < 0.01) & in_contact
    fxy = torch.where(stumble[..., None], fxy * 200.0, fxy)
    cf[:, feet, 2] = fz
    cf[:, feet, 0:2] = fxy
    pen = [b for b in cfg.penalised_contact_indices if b not in cfg.termination_contact_indices]
    hit = (rand(n, len(pen)) < 0.01)[..., None]
    cf[:, pen, :] = 20.0 * randn(n, len(pen), 3) * hit
    base_hit = (rand(n) < 0.003)[:, None]
    for b in cfg.termination_contact_indices:
        cf[:, b, :] = 30.0 * randn(n, 3) * base_hit
    s["contact_forces"] = cf.reshape(n * nb, 3).contiguous()

    # --- rigid bodies: feet under the hips, everything else noise around the root
    rb = torch.zeros(n, nb, 13)
    rb[:, :, 0:3] = s["root_states"][:, None, 0:3] + 0.2 * randn(n, nb, 3)
    rb[:, :, 6] = 1.0
    rb[:, :, 7:13] = 0.5 * randn(n, nb, 6)
    yaw_q = quat.clone()
    yaw_q[:, :2] = 0
    yaw_q = yaw_q / yaw_q.norm(dim=-1, keepdim=True)
    offs = torch.tensor([[0.24, 0.14], [0.24, -0.14], [-0.24, 0.14], [-0.24, -0.14]])
    cz, sz = 1 - 2 * yaw_q[:, 2] ** 2, 2 * yaw_q[:, 2] * yaw_q[:, 3]
    for k, b in enumerate(feet):
        ox, oy = offs[k, 0], offs[k, 1]
        rb[:, b, 0] = s["root_states"][:, 0] + cz * ox - sz * oy
        rb[:, b, 1] = s["root_states"][:, 1] + sz * ox + cz * oy
        rb[:, b, 2] = s["root_states"][:, 2] - 0.35 + 0.1 * rand(n)
    s["rigid_body_states"] = rb.reshape(n * nb, 13).contiguous()

    # --- policy-side buffers
    clipa = lambda t: t.clamp(-cfg.clip_actions, cfg.clip_actions)
    s["actions"] = clipa(0.5 * randn(n, 12))
    s["last_actions"] = clipa(0.5 * randn(n, 12))
    s["last_last_actions"] = clipa(0.5 * randn(n, 12))
    s["last_dof_pos"] = dof_pos + 0.02 * randn(n, 12)
    s["last_dof_vel"] = dof_vel + 0.5 * randn(n, 12)
    s["torques"] = 8.0 * randn(n, 12)
    s["last_torques"] = s["torques"] + 2.0 * randn(n, 12)
    s["last_root_vel"] = 0.5 * randn(n, 6)
    cmd = torch.zeros(n, 4)
    cmd[:, 0] = rand(n) * 2 - 1
    cmd[:, 1] = rand(n) - 0.5
    cmd[:, 2] = rand(n) * 2 - 1
    cmd[:, 3] = (rand(n) * 2 - 1) * math.pi
    small = (cmd[:, :2].norm(dim=1) > 0.2)[:, None]
    cmd[:, :2] *= small
    cmd[:, :2] *= (rand(n) > 0.1)[:, None]          # ~10 % standing-still commands
    s["commands"] = cmd
    s["episode_length_buf"] = torch.randint(0, 1001, (n,), generator=g, dtype=torch.long)
    s["last_contacts"] = rand(n, 4) < 0.5
    s["feet_air_time"] = 0.6 * rand(n, 4) * (rand(n, 4) < 0.7)
    s["Kp_factors"] = 0.9 + 0.2 * rand(n, 1)
    s["Kd_factors"] = 0.9 + 0.2 * rand(n, 1)
    s["motor_strength"] = 0.9 + 0.2 * rand(n, 12)
    s["terrain_levels"] = torch.randint(0, 10, (n,), generator=g, dtype=torch.long)
    dist = torch.zeros(n, nb, 3)
    if step % cfg.disturbance_interval == 0:
        dist[:, 0, :] = (rand(n, 3) * 2 - 1) * 30.0
    s["disturbance"] = dist
    s["obs_buf"] = randn(n, 270).clamp(-cfg.clip_observations, cfg.clip_observations)
    s["privileged_obs_buf"] = randn(n, 51 + len(cfg.measured_points_x) * len(cfg.measured_points_y))
    s["base_lin_vel"] = torch.zeros(n, 3)
    s["base_ang_vel"] = torch.zeros(n, 3)
    s["projected_gravity"] = torch.zeros(n, 3)
    r = len(cfg.episode_sum_names())
    s["episode_sums"] = 0.1 * randn(max(r, 1), n)
    # constants the reference keeps as tensors
    s["default_dof_pos"] = default.clone()
    s["dof_pos_limits"] = torch.stack([torch.from_numpy(tabs["dof_pos_lo"]),
                                       torch.from_numpy(tabs["dof_pos_hi"])], dim=-1)
    s["dof_vel_limits"] = torch.from_numpy(tabs["dof_vel_limits"]).clone()
    s["torque_limits"] = torch.from_numpy(tabs["torque_limits"]).clone()
    return s


def make_noise(n: int, seed: int = 99, n_points: int = 187) -> Dict[str, torch.Tensor]:
    """Pre-drawn U[0,1) tensors standing in for the four `torch.rand_like` draws of one
    post_physics_step (legged_robot.py:451,457 terminal obs; :394,400 obs), in call order."""
    g = torch.Generator().manual_seed(seed)
    return dict(term45=torch.rand(n, 45, generator=g), term187=torch.rand(n, n_points, generator=g),
                obs45=torch.rand(n, 45, generator=g), obs187=torch.rand(n, n_points, generator=g))


def make_reset_targets(cfg: HotPathCfg, state: Dict[str, torch.Tensor], hf: torch.Tensor,
                       seed: int = 4321) -> Dict[str, torch.Tensor]:
    """Deterministic stand-in for what reset_idx's RNG-driven `_reset_dofs`, `_reset_root_states`
    and `_resample_commands` (legged_robot.py:316-320) would write for a reset env: full (N,..)
    tensors; row i is used only if env i resets."""
    n = state["root_states"].shape[0]
    g = torch.Generator().manual_seed(seed)
    rand = lambda *s: torch.rand(*s, generator=g)
    default = state["default_dof_pos"]
    dof_pos = default[None, :] * (0.5 + rand(n, 12))
    dof_vel = (rand(n, 12) * 2 - 1) * 0.1
    root = state["root_states"].clone()
    root[:, 0:2] = torch.stack([rand(n) * cfg.x_size, rand(n) * cfg.y_size], dim=-1) \
        if not cfg.is_plane else rand(n, 2) * 50
    root[:, 2] = _terrain_height_at(cfg, hf, root[:, 0:2]) + 0.5 + 0.05 * rand(n)
    root[:, 3:7] = _quat_from_euler_xyz((rand(n) * 2 - 1) * 0.2, (rand(n) * 2 - 1) * 0.2,
                                        torch.zeros(n))
    root[:, 7:13] = (rand(n, 6) * 2 - 1) * 0.5
    cmd = state["commands"].clone()
    cmd[:, 0] = rand(n) * 2 - 1
    cmd[:, 1] = rand(n) - 0.5
    cmd[:, 3] = (rand(n) * 2 - 1) * math.pi
    return dict(dof_state=torch.stack([dof_pos, dof_vel], dim=-1).reshape(n * 12, 2).contiguous(),
                root_states=root, commands=cmd)


# ----------------------------------------------------------------------------- rollout (GAE)
def make_rollout(n: int, t: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """rewards ~ N(0,0.05^2), values ~ N(0,1), dones ~ Bernoulli(0.002) uint8;
    shapes are the time-major ones of HIMRolloutStorage (him_rollout_storage.py:68-81)."""
    g = torch.Generator().manual_seed(seed)
    return dict(rewards=0.05 * torch.randn(t, n, 1, generator=g),
                values=torch.randn(t, n, 1, generator=g),
                dones=(torch.rand(t, n, 1, generator=g) < 0.002).to(torch.uint8),
                last_values=torch.randn(n, 1, generator=g))


def make_transition(n: int, seed: int = 0, obs_dim: int = 270, priv_dim: int = 238, act_dim: int = 12,
                    reset_frac: float = 0.02, timeout_frac: float = 0.5) -> Dict[str, torch.Tensor]:
    """One synthetic rollout step as the runner sees it after env.step (CPU tensors): the policy's
    transition fields, the env outputs, the reset id list and its terminal rows."""
    g = torch.Generator().manual_seed(9176 + seed)
    randn = lambda *s: torch.randn(*s, generator=g)
    dones = torch.rand(n, generator=g) < reset_frac
    ids = dones.nonzero(as_tuple=False).flatten()
    time_outs = dones & (torch.rand(n, generator=g) < timeout_frac)
    return {
        "obs": randn(n, obs_dim), "critic_obs": randn(n, priv_dim), "privileged_obs": randn(n, priv_dim),
        "termination_ids": ids, "termination_privileged_obs": randn(len(ids), priv_dim),
        "actions": randn(n, act_dim), "rewards": 0.05 * randn(n), "dones": dones, "values": randn(n, 1),
        "time_outs": time_outs, "log_prob": randn(n), "mu": randn(n, act_dim), "sigma": randn(n, act_dim).abs() + 0.1,
    }


def make_filled_storage(n: int, t: int, seed: int = 0, obs_dim: int = 270, priv_dim: int = 238, act_dim: int = 12):
    """Random contents for every (T,N,.) rollout field (CPU tensors, reference attribute names)."""
    g = torch.Generator().manual_seed(5501 + seed)
    r = lambda w: torch.randn(t, n, w, generator=g)
    return {"observations": r(obs_dim), "privileged_observations": r(priv_dim), "next_privileged_observations": r(priv_dim),
            "actions": r(act_dim), "values": r(1), "advantages": r(1), "returns": r(1), "actions_log_prob": r(1),
            "mu": r(act_dim), "sigma": r(act_dim).abs()}


def to_device(d: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
