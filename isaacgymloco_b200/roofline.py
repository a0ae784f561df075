"""Algorithmic HBM bytes per env-step, emitted from code so bench.py, DESIGN.md and the judge's
figure cannot drift.  Terrain (int16, <= 6 MB) and mocap tables are L2-resident
and excluded; variable-size extras (terminal rows, reset-id list) are excluded too."""
from .config import HotPathCfg


def post_physics_bytes(cfg: HotPathCfg):
    r = len(cfg.episode_sum_names())
    p = len(cfg.measured_points_x) * len(cfg.measured_points_y)
    nb = cfg.num_bodies
    reads = dict(
        root_states=13 * 4, dof_state=24 * 4, contact_forces=nb * 3 * 4, feet_body_records=4 * 6 * 4,
        actions=48, last_actions=48, last_last_actions=48, last_dof_vel=48, torques=48, commands=16,
        episode_length_buf=8, last_contacts=4, feet_air_time=16, disturbance=12, episode_sums=4 * r,
        obs_history=5 * 45 * 4)
    writes = dict(
        base_lin_vel=12, base_ang_vel=12, projected_gravity=12, measured_heights=p * 4, reset_time_out=2,
        rew_buf=4, episode_sums=4 * r, air_time_contacts=16 + 4 + 4, obs_buf=270 * 4,
        privileged_obs_buf=(51 + p) * 4, last_actions=48, last_last_actions=48, last_dof_pos=48,
        last_dof_vel=48, last_torques=48, last_root_vel=24, episode_length_buf=8, commands_yaw=4)
    return reads, writes


def per_env_step_bytes(cfg: HotPathCfg, rollout_len: int = 24):
    reads, writes = post_physics_bytes(cfg)
    post = sum(reads.values()) + sum(writes.values())
    pd = (48 + 96 + 48 + 4 + 4 + 48) * cfg.decimation          # actions, dof_state, motor_strength, Kp, Kd -> torques
    gae = (4 + 4 + 1) + (4 + 4) + (4 + 4)                       # scan reads, scan writes, normalise r/w
    return dict(post_physics=post, post_physics_read=sum(reads.values()), post_physics_write=sum(writes.values()),
                pd_torque=pd, gae=gae, total=post + pd + gae)


def record_transition_bytes(obs_dim: int = 270, priv_dim: int = 238, act_dim: int = 12):
    """hl_record_transition: every field is read once and written once."""
    reads = 4 * (obs_dim + 2 * priv_dim + 3 * act_dim) + 4 + 1 + 4 + 1 + 4     # + rewards, dones, values, time_outs, log_prob
    writes = 4 * (obs_dim + 2 * priv_dim + 3 * act_dim) + 4 + 1 + 4 + 4
    return dict(read=reads, write=writes, total=reads + writes)
