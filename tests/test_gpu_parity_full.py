"""-m gpu: parity of the CUDA path on the configurations the bench numbers are quoted on.

* the 65,536-env shard (tile 52 / persistent grid, separate id compaction, one-warp fix-up) against
  the torch oracle run as eager torch on the same B200;
* a 24-step rollout (the loop bench.py times: 4 x PD torque, fused step, ids + terminal rows,
  fix-up with the in-kernel buffer resets) against the oracle stepped 24 times, fresh synthetic
  PhysX state every step;
* two env shards with env_id_offset = 0 / N against one 2N-env run (index-dependent semantics:
  stumble slices legged_robot.py:1597-1598, Philox streams keyed by the global env id);
* _get_base_heights() values (legged_robot.py:1357-1398);
* antipodal / identical / near-identical quaternion pairs through hl_amp_frame_blend
  (rsl_rl/rsl_rl/utils/utils.py:153-186).

Bit-exact: ids, flags, counters, height cells.  Floats: rel 1e-5 (+ atol 2e-6), written in
tests/gpu_helpers.py.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PHYSX = ("root_states", "dof_state", "contact_forces", "rigid_body_states")


def _oracle(cfg, state, hf, device="cuda"):
    from oracle import torch_oracle as O
    from isaacgymloco_b200 import synthetic as S
    return O.OracleEnv(cfg, S.to_device(state, device), hf.to(device))


@pytest.mark.parametrize("task", ["flat", "stairs"])
def test_fused_step_65536_vs_oracle(task):
    """BASELINE.json configs[4]'s shard size, the one every perf number is quoted on: fused kernel
    + hl_select_and_terminal + torch reset_idx + fix-up == oracle (eager torch on the same GPU),
    pre-drawn noise (Philox off)."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 65536
    cfg = C.aliengo(task, num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=1234)
    noise = S.make_noise(n, seed=4)
    targets = S.make_reset_targets(cfg, state, hf, seed=5)
    oenv = _oracle(cfg, state, hf)
    oids, oterm, oamp = oenv.post_physics_step(S.to_device(noise, "cuda"), S.to_device(targets, "cuda"))
    want = oenv.snapshot()
    del oenv
    torch.cuda.empty_cache()
    env = make_env(cfg, state, hf, targets, noise)
    assert not env.single_launch
    ids, term_obs, term_amp = env.post_physics_step()
    assert len(oids) > 500
    assert_equal(ids, oids, "env_ids")
    assert_close(term_obs, oterm, "termination_privileged_obs")
    assert_close(term_amp, oamp, "terminal_amp_states")
    compare_snapshots(env.snapshot(), want)


@pytest.mark.parametrize("task,n", [("flat", 4096), ("stairs", 16384), ("flat", 65536)])
def test_rollout_24_steps_vs_oracle(task, n):
    """The loop bench.py times, 24 env-steps with fresh PhysX tensors and actions every step:
    history roll over more than 6 steps, feet_air_time, episode sums, episode_length_buf / time-outs,
    last_* chains and the reset bookkeeping stay equal to the oracle at every step."""
    from gpu_helpers import assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    t_len = 24
    cfg = C.aliengo(task, num_envs=n)
    hf = S.make_terrain(cfg, seed=2)
    state = S.make_state(cfg, n, hf, seed=300)
    oenv = _oracle(cfg, state, hf)
    env = make_env(cfg, state, hf)
    total_resets = 0
    check_every = 1 if n <= 16384 else 6
    for step in range(t_len):
        fresh = S.make_state(cfg, n, hf, seed=300, step=step + 1)
        noise = S.make_noise(n, seed=900 + step)
        for k in PHYSX + ("actions", "disturbance"):
            getattr(env, k).view(-1).copy_(fresh[k].view(-1).cuda())
            getattr(oenv, k).view(-1).copy_(fresh[k].view(-1).cuda())
        env.set_noise_tensors(**noise)
        delayed = env.actions.view(n, 1, 12).repeat(1, 4, 1).contiguous()
        for k in range(4):
            env._compute_torques_into(delayed[:, k], env.torques)
            oenv.torques = oenv._compute_torques(delayed[:, k])
        oids, _, _ = oenv.post_physics_step(S.to_device(noise, "cuda"), {})     # {}: RNG-free reset only
        env.fused_pre_reset()
        env.fused_post_reset(with_reset_zero=True)
        env.common_step_counter += 1
        k = int(env._n_reset.item())
        total_resets += k
        assert_equal(env._reset_ids[:k], oids, f"env_ids at step {step}")
        if step % check_every == check_every - 1 or step == t_len - 1:
            compare_snapshots(env.snapshot(), oenv.snapshot())
    assert total_resets > t_len            # resets happened all along the rollout


def test_philox_streams_differ_across_steps_and_match_across_launch_modes():
    """Throughput mode: the Philox offset advances with common_step_counter (fresh noise every step,
    uncorrelated with the previous step's), and the stream depends on (seed, step, global env id)
    only -- not on the tile size or launch mode."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 8192
    cfg = C.aliengo("flat", num_envs=n)
    cfg0 = C.aliengo("flat", num_envs=n, add_noise=False)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=5)
    noisy, clean = make_env(cfg, state, hf), make_env(cfg0, state, hf)
    nv = torch.tensor(cfg.noise_scale_vec())
    full = torch.cat([nv[:45], torch.zeros(6), nv[45:]])
    us = []
    for step in range(3):
        for e in (noisy, clean):
            e.fused_pre_reset()
            e.fused_post_reset(with_reset_zero=True)
            e.common_step_counter += 1
        keep = ~noisy.reset_buf.cpu()
        d = (noisy.privileged_obs_buf - clean.privileged_obs_buf).cpu()
        us.append((d[:, full != 0] / full[full != 0])[keep])
    m = min(u.shape[0] for u in us)
    for a in range(3):
        for b in range(a + 1, 3):
            assert not torch.equal(us[a][:m], us[b][:m])
            corr = torch.corrcoef(torch.stack([us[a][:m].flatten(), us[b][:m].flatten()]))[0, 1]
            assert abs(float(corr)) < 5e-3, (a, b, float(corr))
    # same seed / step / env ids through the other launch mode: identical draws
    other = make_env(cfg, state, hf)
    other.single_launch = not other.single_launch
    other.refresh_buffers()
    ref = make_env(cfg, state, hf)
    for e in (other, ref):
        e.fused_pre_reset()
    assert torch.equal(other.privileged_obs_buf, ref.privileged_obs_buf)
    assert torch.equal(other._term_priv[:int(other._n_reset.item())], ref._term_priv[:int(ref._n_reset.item())])


@pytest.mark.parametrize("n,split", [(16384, 6144), (65536, 24576)])
def test_two_shards_equal_one_big_env(n, split):
    """Env-sharded data parallel: shards [0,split) and [split,n) run with
    env_id_offset 0 / split must reproduce the single n-env run row for row -- stumble index slices
    (LR:1597-1598) and the Philox streams are keyed by the GLOBAL env id.  In-kernel Philox noise,
    stairs reward set; the shard boundary lies inside the stairs-up stumble slice."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    cfg = C.aliengo("stairs", num_envs=n)
    hf = S.make_terrain(cfg, seed=3)
    state = S.make_state(cfg, n, hf, seed=77)
    state["terrain_levels"][:] = torch.randint(4, 10, (n,), generator=torch.Generator().manual_seed(1))
    sr = cfg.stumble_ranges()
    assert sr["stairsup_start"] < split < sr["stairsup_end"], "the stumble slice must straddle the shard boundary"
    big = make_env(cfg, state, hf)
    shards, bounds = [], [(0, split), (split, n)]
    for lo, hi in bounds:
        cfg_r = C.aliengo("stairs", num_envs=n, env_id_offset=lo)
        st_r = {}
        for k, v in state.items():
            if k == "episode_sums":
                st_r[k] = v[:, lo:hi].contiguous()
            elif isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] in (n, n * 12, n * cfg.num_bodies):
                per = v.shape[0] // n
                st_r[k] = v[lo * per:hi * per].contiguous()
            else:
                st_r[k] = v
        shards.append(make_env(cfg_r, st_r, hf))
    for step in range(2):
        for e in [big] + shards:
            e.fused_pre_reset()
            e.fused_post_reset(with_reset_zero=True)
            e.common_step_counter += 1
        kb = int(big._n_reset.item())
        ks = [int(s._n_reset.item()) for s in shards]
        assert kb == sum(ks) and kb > 0
        merged = torch.cat([shards[0]._reset_ids[:ks[0]], shards[1]._reset_ids[:ks[1]] + split])
        assert torch.equal(big._reset_ids[:kb], merged)
        assert torch.equal(big._term_priv[:kb], torch.cat([s._term_priv[:k] for s, k in zip(shards, ks)]))
        sb = big.snapshot()
        ss = [s.snapshot() for s in shards]
        for key in sb:
            cat_dim = 1 if key == "episode_sums" else 0
            got = torch.cat([x[key] for x in ss], dim=cat_dim)
            assert torch.equal(sb[key], got), (step, key)
    stum = big.cfg_hot.episode_sum_names().index("feet_stumble")
    term = torch.zeros(n)
    # the stumble term alone (scale -1.0 * dt): nonzero only inside the slice, on both sides of the boundary
    oenv_terms = (state["contact_forces"].view(n, -1, 3)[:, cfg.feet_indices, :2].norm(dim=-1)
                  > 5 * state["contact_forces"].view(n, -1, 3)[:, cfg.feet_indices, 2].abs()).any(dim=1)
    term[sr["stairsup_start"]:sr["stairsup_end"]] = oenv_terms[sr["stairsup_start"]:sr["stairsup_end"]].float()
    assert term[:split].sum() > 0 and term[split:].sum() > 0, "stumbling envs on both sides of the boundary"
    del stum


@pytest.mark.parametrize("task,oracle_dev", [("flat", "cuda"), ("stairs", "cuda"), ("stairs", "cpu")])
def test_get_base_heights_values(task, oracle_dev):
    """_get_base_heights() (LR:1357-1398): 63-point scan, mean(root_z - h)."""
    from gpu_helpers import assert_close, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 4096 + 3
    cfg = C.aliengo(task, num_envs=n,
                    index_math=C.INDEX_MATH_TORCH_CPU if oracle_dev == "cpu" else C.INDEX_MATH_TORCH_CUDA)
    hf = S.make_terrain(cfg, seed=9)
    state = S.make_state(cfg, n, hf, seed=61)
    want = _oracle(cfg, state, hf, oracle_dev)._get_base_heights()
    env = make_env(cfg, state, hf)
    got = env._get_base_heights()
    assert got.shape == (n,)
    assert float(want.std()) > 0.01
    assert_close(got, want, "base heights")


def test_amp_blend_antipodal_and_identical_quaternions():
    """quaternion_slerp's special cases (utils.py:153-186) through hl_amp_frame_blend: consecutive
    frames whose root quaternions are identical (|d| = 1 -> q0), antipodal (d = -1 -> q0 after the
    shortest-path flip), nearly identical (angle tiny but above eps: the 1/angle scaling), opposite
    hemisphere (d < 0: q1 negated) and orthogonal (d = 0), at blend = 0, 1, within isclose() of 0 / 1
    and in between."""
    from oracle import torch_oracle as O
    from isaacgymloco_b200.motion_loader import AMPLoader
    rng = np.random.default_rng(3)

    def unit(q):
        return q / np.linalg.norm(q, axis=-1, keepdims=True)

    q = unit(rng.normal(size=(4,)))
    r = unit(rng.normal(size=(4,)))
    tiny = unit(q + 1e-4 * rng.normal(size=4))
    tinier = unit(q + 3e-7 * rng.normal(size=4))
    ortho = unit(np.array([-q[1], q[0], -q[3], q[2]]))
    quats = [q, q, -q, tiny, q, tinier, -tinier, r, -r, ortho, q, -ortho, ortho]
    nf = len(quats)
    frames = rng.normal(size=(nf, 61))
    frames[:, 3:7] = np.stack(quats)
    dur = 0.02
    tabs = dict(frames=[frames.astype(np.float32)], frame_durations=[dur], weights=[1.0], names=["adv"])
    ld = AMPLoader("cuda:0", 0.02, clip_tables=tabs)
    tab = O.OracleMotionTable([torch.from_numpy(frames.astype(np.float32)[:, :49])], [dur], [1.0], 0.02)
    length = tab.trajectory_lens[0]
    fr = np.array([0.0, 1e-9, 4e-6, 0.25, 0.5, 0.9, 1.0 - 4e-6, 1.0 - 1e-9])
    pn = (np.arange(nf - 1)[:, None] + fr[None, :]).reshape(-1)          # target p*n values
    times = pn * length / nf
    ok = np.ceil(times / length * nf) <= nf - 1
    times = times[ok]
    idx = np.zeros(len(times), dtype=np.int64)
    want, lo, hi = tab.get_full_frame_at_time_batch(idx, times)
    got, glo, ghi = ld.get_full_frame_at_time_batch(idx, times, return_indices=True)
    np.testing.assert_array_equal(glo.cpu().numpy(), lo.astype(np.int32))
    np.testing.assert_array_equal(ghi.cpu().numpy(), hi.astype(np.int32))
    assert set(np.unique(lo)) >= set(range(nf - 2)), "every quaternion pair is exercised"
    g, w = got.cpu().numpy(), want.numpy()
    lerp_cols = [c for c in range(49) if not 3 <= c < 7]
    np.testing.assert_array_equal(g[:, lerp_cols], w[:, lerp_cols])
    np.testing.assert_allclose(g[:, 3:7], w[:, 3:7], rtol=1e-5, atol=1e-6, equal_nan=True)
    # the special-case rows are exact copies of q0 in the reference: they must be exact here too
    d = np.sum(frames[lo, 3:7].astype(np.float32) * frames[hi, 3:7].astype(np.float32), axis=-1)
    special = np.abs(np.abs(d) - 1.0) < np.finfo(float).eps * 4.0
    assert special.sum() >= 8
    np.testing.assert_array_equal(g[special, 3:7], w[special, 3:7])
