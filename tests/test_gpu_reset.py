"""-m gpu: reset_idx inside the kernel chain and the drop-in wiring of step() /
post_physics_step() (legged_robot.py:122-247,288-361,607-656).

* the kernel's re-draws, fed the same uniform table, against fixtures minted from the reference's own reset_idx
  (torch_rand_float / rand_like / randint_like patched to read that table in call order) and against the oracle;
* several steps of post_physics_step with the in-kernel reset + pre-step command resampling against the oracle;
* throughput mode (Philox stream 3): every re-drawn quantity is uniform on the reference's range
  (Kolmogorov-Smirnov + moments), independent across columns and envs -- statistical, not bitwise, parity with
  torch's generator;
* a FakeGym that records the call order and mutates the state between substeps: step() issues the reference's
  gym.* calls in the reference's order, pushes / disturbances / command resampling fire on the reference's steps,
  the AMP cfg returns the 8-tuple;
* post_physics_step_device() is CUDA-graph capturable (no host sync, no allocation).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _attach_terrain(env, ter, cfg):
    dev = env.device
    env.custom_origins = True
    env.env_origins = ter["env_origins"].to(dev).contiguous()
    env.terrain_origins = ter["origins"].to(dev).contiguous()
    env.terrain_types = ter["types"].to(dev).contiguous()
    env.max_terrain_level = cfg.num_rows
    env.refresh_buffers()


def _oracle_terrain(ter, cfg, dev):
    return dict(origins=ter["origins"].to(dev), types=ter["types"].to(dev), max_level=cfg.num_rows, env_length=cfg.terrain_length,
                max_episode_length_s=cfg.episode_length_s)


@pytest.mark.parametrize("name", ["flat", "stairs"])
def test_reset_idx_kernel_vs_reference_golden(name):
    from gpu_helpers import assert_close, assert_equal
    from isaacgymloco_b200.legged_robot import FusedLeggedRobot
    from oracle.make_goldens import RESET_CASES, reset_inputs
    case = [c for c in RESET_CASES if c["name"] == name][0]
    gold = load_golden(f"reset_{name}.npz")
    cfg, hf, state, u, ter, ids = reset_inputs(case)
    env = FusedLeggedRobot(cfg, state, hf, device="cuda:0")
    _attach_terrain(env, ter, cfg)
    env.set_reset_uniforms(u)
    env.reset_idx_device(ids.cuda())
    assert_equal(env.terrain_levels, gold["terrain_levels"], "terrain_levels")
    for k in ("root_states", "dof_state", "commands", "Kp_factors", "Kd_factors", "motor_strength_factors", "env_origins"):
        assert_close(getattr(env, k).reshape(gold[k].shape), gold[k], k, rtol=1e-6, atol=1e-6)
    for k in ("last_actions", "last_last_actions", "last_dof_pos", "last_dof_vel", "last_torques", "feet_air_time"):
        assert_equal(getattr(env, k), gold[k], k)
    assert_equal(env.episode_length_buf, gold["episode_length_buf"], "episode_length_buf")
    assert_equal(env.reset_buf[ids.cuda()], np.ones(len(ids), dtype=bool), "reset_buf")
    # the same through the reference's method names (individual hooks + bookkeeping) on a fresh env
    env2 = FusedLeggedRobot(cfg, state, hf, device="cuda:0")
    _attach_terrain(env2, ter, cfg)
    env2.set_reset_uniforms(u)
    env2.reset_idx(ids.cuda())
    for k in ("root_states", "dof_state", "commands", "Kp_factors", "Kd_factors", "motor_strength_factors", "env_origins",
              "terrain_levels", "last_actions", "feet_air_time", "episode_length_buf"):
        assert torch.equal(getattr(env, k), getattr(env2, k)), k
    for nm, val in zip(gold["extras_names"], gold["extras_vals"]):
        assert_close(env2.extras["episode"][str(nm)], val, str(nm), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("task,n", [("flat", 4096), ("stairs", 16384)])
def test_post_physics_step_with_kernel_reset_vs_oracle(task, n):
    """Six steps of the public post_physics_step (fused step, ids + terminal rows, in-kernel reset_idx + fix-up in
    one launch, pre-step command resampling on the 500-step mark) against the oracle driven with the same
    uniform tables and noise."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots
    from isaacgymloco_b200 import config as C, synthetic as S
    from isaacgymloco_b200.legged_robot import FusedLeggedRobot
    from oracle import torch_oracle as O
    from oracle.make_goldens import reset_inputs
    cfg, hf, state, _, ter, _ = reset_inputs(dict(task=task, n=n, seed=71))
    cfg.reset.push_robots = False
    cfg.reset.disturbance = False
    state["episode_length_buf"][: n // 16] = 498          # these hit the 500-step resampling mark on steps 2, 3
    state["episode_length_buf"][n // 16: n // 8] = 499
    dev = "cuda:0"
    env = FusedLeggedRobot(cfg, state, hf, device=dev)
    _attach_terrain(env, ter, cfg)
    assert env._kernel_reset_ok() and env.resample_interval == 500
    oenv = O.OracleEnv(cfg, S.to_device(state, dev), hf.to(dev))
    oenv.env_origins = ter["env_origins"].to(dev).clone()
    oter = _oracle_terrain(ter, cfg, dev)
    total = 0
    for step in range(6):
        g = torch.Generator().manual_seed(500 + step)
        u = torch.rand(n, C.RESET_NU, generator=g).to(dev)
        noise = S.make_noise(n, seed=600 + step)
        fresh = S.make_state(cfg, n, hf, seed=71, step=step + 1)
        for k in ("contact_forces", "rigid_body_states", "actions"):        # (root / dof rows carry the reset draws forward)
            getattr(env, k).view(-1).copy_(fresh[k].view(-1).to(dev))
            getattr(oenv, k).view(-1).copy_(fresh[k].view(-1).to(dev))
        env.set_noise_tensors(**noise)
        env.set_reset_uniforms(u)
        # ---- oracle: LR:612-613 resampling on the mark, the step, reset_idx with the same uniforms
        mark = ((oenv.episode_length_buf + 1) % 500 == 0).nonzero(as_tuple=False).flatten()
        oenv.resample_commands(mark, u, cfg.reset)
        nz = S.to_device(noise, dev)
        oids, oterm, oamp = oenv.pre_reset(nz)
        oenv.reset_idx_draw(oids, u, custom_origins=True, terrain=oter)
        oenv.apply_reset(oids, {})
        oenv.post_reset(nz)
        # ---- the public call
        ids, term_obs, term_amp = env.post_physics_step()
        total += len(ids)
        assert_equal(ids, oids, f"env_ids step {step}")
        assert_close(term_obs, oterm, "termination_privileged_obs")
        assert_close(term_amp, oamp, "terminal_amp_states")
        compare_snapshots(env.snapshot(), oenv.snapshot())
        assert_equal(env.terrain_levels, oenv.terrain_levels, "terrain_levels")
        for k in ("root_states", "dof_state", "env_origins", "Kp_factors", "Kd_factors", "motor_strength_factors"):
            assert_close(getattr(env, k).reshape(getattr(oenv, k).shape), getattr(oenv, k), k, rtol=1e-6, atol=1e-6)
        if step >= 1:
            assert len(mark) > 0 or step > 2
    assert total > 6 and env.common_step_counter == 6


def test_reset_draws_philox_statistics():
    """Throughput mode: the in-kernel Philox draws of reset_idx are uniform on the reference's ranges."""
    from scipy import stats
    from isaacgymloco_b200 import config as C, synthetic as S
    from isaacgymloco_b200.legged_robot import FusedLeggedRobot
    n = 32768
    cfg = C.aliengo("flat", num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=3)
    env = FusedLeggedRobot(cfg, state, hf, device="cuda:0", seed=123)
    env.custom_origins = True
    ids = torch.arange(n, device="cuda")
    env.reset_idx_device(ids)
    R = cfg.reset
    t = cfg.dof_tables()
    dof = env.dof_state.view(n, 12, 2).cpu().numpy()
    cols = {}
    for d in (1, 2, 7, 11):                      # thigh / calf dofs (hips have default 0: ratio not observable)
        cols[f"dof_pos{d}"] = ((dof[:, d, 0] / t["default_dof_pos"][d]), R.dof_init_pos_ratio_range)
    for d in (0, 5):
        cols[f"dof_vel{d}"] = (dof[:, d, 1], R.dof_init_vel_range)
    root = env.root_states.cpu().numpy()
    cols["x"] = (root[:, 0] - R.base_init_state[0], R.base_init_pos_range["x"])
    cols["z"] = (root[:, 2] - R.base_init_state[2], R.base_init_pos_range["z"])
    for k, ax in enumerate(("x", "y", "z", "roll", "pitch", "yaw")):
        cols["vel_" + ax] = (root[:, 7 + k], R.base_init_vel_range[ax])
    cols["kp"] = (env.Kp_factors.cpu().numpy()[:, 0], R.kp_range)
    cols["kd"] = (env.Kd_factors.cpu().numpy()[:, 0], R.kd_range)
    cols["ms"] = (env.motor_strength_factors.cpu().numpy()[:, 0], R.motor_strength_range)
    cmd = env.commands.cpu().numpy()
    cols["heading"] = (cmd[:, 3], R.heading)
    us = {}
    for name, (x, (lo, hi)) in cols.items():
        uu = (x - lo) / (hi - lo)
        assert uu.min() >= -1e-6 and uu.max() <= 1 + 1e-6, name
        p = stats.kstest(np.clip(uu, 0, 1), "uniform").pvalue
        assert p > 1e-4, (name, p)
        assert abs(uu.mean() - 0.5) < 0.01 and abs(uu.var() - 1 / 12) < 0.005, name
        us[name] = uu
    names = list(us)
    m = np.corrcoef(np.stack([us[k] for k in names]))
    assert np.abs(m - np.eye(len(names))).max() < 0.03
    # roll / pitch from the quaternion (yaw range is [0, 0]): small-angle check of quat_from_euler_xyz
    q = root[:, 3:7]
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-5)
    roll = np.arctan2(2 * (q[:, 3] * q[:, 0] + q[:, 1] * q[:, 2]), 1 - 2 * (q[:, 0] ** 2 + q[:, 1] ** 2))
    assert roll.min() >= -0.2 - 1e-4 and roll.max() <= 0.2 + 1e-4 and abs(roll.var() - 0.4 ** 2 / 12) < 1e-3
    # commands: LR:649-656 invariants
    nrm = np.linalg.norm(cmd[:, :2], axis=1)
    assert ((nrm == 0) | (nrm > 0.2)).all()
    hi_ids = np.arange(n) < 0.2 * n
    assert (np.abs(cmd[~hi_ids, 0]) <= 1.0).all() and (cmd[hi_ids & (np.abs(cmd[:, 0]) >= 1.0), 1] == 0).all()
    # a second step draws a fresh, uncorrelated set
    first = env.root_states[:, 7].clone()
    env.common_step_counter += 1
    env.reset_idx_device(ids)
    c = np.corrcoef(first.cpu().numpy(), env.root_states[:, 7].cpu().numpy())[0, 1]
    assert abs(c) < 0.03


class FakeGym:
    """Records every gym.* call; `simulate` advances the dof state like a physics step would."""

    def __init__(self, env_ref):
        self.calls, self.env = [], env_ref

    def __getattr__(self, name):
        def fn(*a, **k):
            self.calls.append(name)
            if name == "simulate":
                e = self.env()
                e.dof_state.view(e.num_envs, 12, 2)[..., 0] += 0.001 * e.torques        # "physics"
                e.dof_state.view(e.num_envs, 12, 2)[..., 1] += 0.01 * e.torques
        return fn


def test_step_drives_the_gym_like_the_reference():
    """step(): per substep set_dof_actuation_force_tensor, simulate, fetch_results, refresh_dof_state_tensor
    (LR:148-152), then the four refresh_* of LR:187-190; pushes every push_interval steps (root velocities the
    rewards see), disturbances every 8, the indexed PhysX setters after a reset; 8-tuple under USING_AMP."""
    import weakref
    from isaacgymloco_b200 import config as C, synthetic as S
    from isaacgymloco_b200.legged_robot import FusedLeggedRobot
    n = 2048
    cfg = C.aliengo("amp", num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=9)
    env = FusedLeggedRobot(cfg, state, hf, device="cuda:0")
    env.custom_origins = True
    env.push_interval = 3                        # push on steps 3, 6, ...
    env.gym, env.sim = FakeGym(weakref.ref(env)), object()
    pushes, dists = [], []
    orig_push, orig_dist = env._push_robots, env._disturbance_robots
    env._push_robots = lambda: (pushes.append(env.common_step_counter), orig_push())
    env._disturbance_robots = lambda: (dists.append(env.common_step_counter), orig_dist())
    ref = FusedLeggedRobot(cfg, state, hf, device="cuda:0")     # same env without a gym, driven by physics_step_fn
    ref.custom_origins = True
    ref.push_interval = 3

    def phys(e, k):
        e.dof_state.view(e.num_envs, 12, 2)[..., 0] += 0.001 * e.torques
        e.dof_state.view(e.num_envs, 12, 2)[..., 1] += 0.01 * e.torques
    ref.physics_step_fn = phys
    cfg.reset.delay = False
    torch.manual_seed(0)
    acts = [0.3 * torch.randn(n, 12, device="cuda") for _ in range(8)]
    for i, a in enumerate(acts):
        env.gym.calls.clear()
        torch.manual_seed(100 + i)
        out = env.step(a)
        torch.manual_seed(100 + i)
        out_ref = ref.step(a)
        assert len(out) == 8 and out[7].shape[1] == 30 and torch.equal(out[7], out_ref[7])
        calls = env.gym.calls
        sub = ["set_dof_actuation_force_tensor", "simulate", "fetch_results", "refresh_dof_state_tensor"]
        assert calls[:16] == sub * 4, calls[:16]
        assert calls[16:20] == ["refresh_actor_root_state_tensor", "refresh_net_contact_force_tensor", "refresh_force_sensor_tensor",
                                "refresh_rigid_body_state_tensor"]
        rest = calls[20:]
        if env.common_step_counter % 8 == 0:
            assert "apply_rigid_body_force_tensors" in rest
        if env.common_step_counter % 3 == 0:
            assert "set_actor_root_state_tensor" in rest
        if len(out[5]):
            assert rest[-2:] == ["set_dof_state_tensor_indexed", "set_actor_root_state_tensor_indexed"]
        for x, y in zip(out[:4], out_ref[:4]):
            assert torch.equal(x, y)
        assert torch.equal(out[5], out_ref[5]) and torch.equal(out[6], out_ref[6])
    assert pushes == [3, 6] and dists == [8] and env.common_step_counter == 8


def test_post_physics_step_device_is_graph_capturable():
    """No host sync and no allocation inside the chain: capture one step, replay it, compare with eager."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    from isaacgymloco_b200.legged_robot import FusedLeggedRobot
    n = 4096
    cfg = C.aliengo("flat", num_envs=n)
    cfg.reset.push_robots = False
    cfg.reset.disturbance = False
    cfg.reset.commands_curriculum = False
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=5)
    a = FusedLeggedRobot(cfg, state, hf, device="cuda:0", seed=7)
    b = FusedLeggedRobot(cfg, state, hf, device="cuda:0", seed=7)
    for e in (a, b):
        e.custom_origins = True
        e.post_physics_step()                     # warm-up (first step clips the history)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            a.post_physics_step_device()
    torch.cuda.current_stream().wait_stream(s)
    # the capture did not execute: rewind the host counter it advanced, then replay
    a.common_step_counter -= 1
    a._buffers()
    g.replay()
    a.common_step_counter += 1
    b.post_physics_step()
    torch.cuda.synchronize()
    sa, sb = a.snapshot(), b.snapshot()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(a.root_states, b.root_states) and int(a._n_reset.item()) == int(b._n_reset.item()) > 0


def test_packed_foot_records_equal_rigid_body_states():
    """HlEnvBuffers.foot_records (N,4,13), the input a PCIe-fed host should ship, gives the same step as reading
    rigid_body_states[:, feet_indices] (LR:203-204) -- fused kernel, fix-up and the staged methods."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import _lib as L
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 4096 + 4
    cfg = C.aliengo("stairs", num_envs=n)
    hf = S.make_terrain(cfg, seed=4)
    state = S.make_state(cfg, n, hf, seed=17)
    a, b = make_env(cfg, state, hf), make_env(cfg, state, hf)
    b.foot_records = b.rigid_body_states.view(n, cfg.num_bodies, 13)[:, cfg.feet_indices, :].contiguous()
    b.rigid_body_states = torch.full_like(b.rigid_body_states, float("nan"))     # must not be read any more
    b.refresh_buffers()
    for e in (a, b):
        e.fused_pre_reset()
        e.fused_post_reset(with_reset_zero=True)
    sa, sb = a.snapshot(), b.snapshot()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    for e in (a, b):
        e._stages(L.ST_FRAME | L.ST_CONTACTS | L.ST_REWARD)
    assert torch.equal(a.rew_buf, b.rew_buf) and torch.equal(a.feet_pos, b.feet_pos)
