"""-m gpu: the CUDA post-physics path (through the C ABI) against (1) the fixtures minted from the
real reference, (2) the torch oracle on CPU and on the same GPU, at sizes up to BASELINE.json's
65,536 envs.  Bit-exact: reset ids, termination/time-out flags, contact flags, height-field cell
indices (hence measured_heights).  Floats: rel 1e-5 (+ atol 2e-6 near zero)."""
import numpy as np
import pytest
import torch

from conftest import build_case_inputs, load_golden

pytestmark = pytest.mark.gpu


def _cases():
    from oracle.make_goldens import ENV_CASES
    return ENV_CASES


@pytest.mark.parametrize("case", _cases(), ids=[c["name"] for c in _cases()])
def test_fused_step_vs_reference_golden(case):
    """Fused kernel + id compaction + terminal rows + (deterministic) reset + fix-up == the real
    reference's post_physics_step + obs clip on the same inputs (fixtures minted on CPU => the
    kernel runs with the torch-CPU index arithmetic)."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    gold = load_golden(f"env_{case['name']}.npz")
    cfg, hf, state, noise, targets, delayed, chk = build_case_inputs(case)
    assert chk == gold["input_checksum"]
    env = make_env(cfg, state, hf, targets, noise, no_reset=case.get("no_reset", False))
    tq = torch.stack([env._compute_torques(delayed.cuda()[:, k]) for k in range(4)], dim=1)
    assert_close(tq, gold["torques4"], "torques")
    if not cfg.is_plane:
        idx = env.get_height_indices().cpu().numpy()
        assert_equal(idx[..., 0], gold["px"].astype(np.int32), "px")
        assert_equal(idx[..., 1], gold["py"].astype(np.int32), "py")
    ids, term_obs, term_amp = env.post_physics_step()
    assert ids.dtype == torch.int64
    assert_equal(ids, gold["env_ids"], "env_ids")
    assert_close(term_obs, gold["term_obs"], "termination_privileged_obs")
    assert_close(term_amp, gold["term_amp"], "terminal_amp_states")
    compare_snapshots(env.snapshot(), {k: gold[k] for k in gold})
    if "extras_names" in gold and len(gold["env_ids"]):
        for name, val in zip(gold["extras_names"], gold["extras_vals"]):
            assert_close(env.extras["episode"][str(name)], val, str(name), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", _cases(), ids=[c["name"] for c in _cases()])
def test_individual_methods_vs_reference_golden(case):
    """The same step driven through the individual drop-in methods, in the reference's order."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    gold = load_golden(f"env_{case['name']}.npz")
    cfg, hf, state, noise, targets, delayed, _ = build_case_inputs(case)
    env = make_env(cfg, state, hf, targets, noise, no_reset=case.get("no_reset", False))
    env.episode_length_buf += 1                              # LR:193
    env.update_base_frame()                                  # LR:197-209
    env.update_heading_command()                             # LR:616-620
    env.measured_heights = env._get_heights()                # LR:623-624
    env.check_termination()                                  # LR:219
    env.compute_reward()                                     # LR:222
    ids = env.reset_buf.nonzero(as_tuple=False).flatten()    # LR:225
    term_obs = env.compute_termination_observations(ids).clone()   # LR:227
    term_amp = env.get_amp_observations()[ids]               # LR:228
    env.reset_idx(ids)                                       # LR:229
    if len(ids) and not case.get("no_reset", False):
        env.measured_heights = env._get_heights()            # LR:332-333
    env.compute_observations()                               # LR:232
    # LR:235-241 + LR:167-171
    env.disturbance[:] = 0.0
    env.last_last_actions[:] = env.last_actions[:]
    env.last_actions[:] = env.actions[:]
    env.last_dof_pos[:] = env.dof_pos[:]
    env.last_dof_vel[:] = env.dof_vel[:]
    env.last_torques[:] = env.torques[:]
    env.last_root_vel[:] = env.root_states[:, 7:13]
    env.obs_buf.clamp_(-cfg.clip_observations, cfg.clip_observations)
    env.privileged_obs_buf.clamp_(-cfg.clip_observations, cfg.clip_observations)
    assert_equal(ids, gold["env_ids"], "env_ids")
    assert_close(term_obs, gold["term_obs"], "termination_privileged_obs")
    assert_close(term_amp, gold["term_amp"], "terminal_amp_states")
    compare_snapshots(env.snapshot(), {k: gold[k] for k in gold})
    bh = env._get_base_heights()
    assert bh.shape == (env.num_envs,)


def _oracle_step(cfg, state, hf, noise, targets, device):
    from oracle import torch_oracle as O
    from isaacgymloco_b200 import synthetic as S
    env = O.OracleEnv(cfg, S.to_device(state, device), hf.to(device))
    nz = S.to_device(noise, device)
    ids, term_obs, term_amp = env.post_physics_step(nz, S.to_device(targets, device) if targets else None)
    return env, ids.cpu(), term_obs.cpu(), term_amp.cpu()


@pytest.mark.parametrize("task,n,oracle_dev", [("flat", 4096, "cpu"), ("stairs", 16384, "cpu"),
                                               ("flat", 4096, "cuda"), ("stairs", 16384, "cuda"),
                                               ("recover", 2048, "cuda"), ("amp", 2048, "cuda")])
def test_fused_step_vs_oracle(task, n, oracle_dev):
    """BASELINE.json configs[1] (flat, 4096) and configs[2] (stairs, 16384) against the torch
    oracle run on CPU (torch-CPU index arithmetic) and as eager torch on the same B200 (the
    device the reference runs on in production; torch-CUDA index arithmetic)."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    cfg = C.aliengo(task, num_envs=n,
                    index_math=C.INDEX_MATH_TORCH_CPU if oracle_dev == "cpu" else C.INDEX_MATH_TORCH_CUDA)
    hf = S.make_terrain(cfg, seed=5)
    state = S.make_state(cfg, n, hf, seed=77)
    noise = S.make_noise(n, seed=78)
    targets = S.make_reset_targets(cfg, state, hf, seed=79)
    oenv, oids, oterm, oamp = _oracle_step(cfg, state, hf, noise, targets, oracle_dev)
    env = make_env(cfg, state, hf, targets, noise)
    ids, term_obs, term_amp = env.post_physics_step()
    assert len(oids) > 0
    assert_equal(ids, oids, "env_ids")
    assert_close(term_obs, oterm, "termination_privileged_obs")
    assert_close(term_amp, oamp, "terminal_amp_states")
    compare_snapshots(env.snapshot(), oenv.snapshot())


@pytest.mark.parametrize("oracle_dev", ["cpu", "cuda"])
def test_height_indices_bit_exact(oracle_dev):
    """3.06 M scan points (16,384 envs x 187) on the 1300x2300 stairs field, including envs placed
    outside the terrain (clipped cells): every (px,py) equals eager torch's on that device."""
    from gpu_helpers import assert_equal, make_env
    from oracle import torch_oracle as O
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 16384
    cfg = C.aliengo("stairs", num_envs=n,
                    index_math=C.INDEX_MATH_TORCH_CPU if oracle_dev == "cpu" else C.INDEX_MATH_TORCH_CUDA)
    hf = S.make_terrain(cfg, seed=6)
    state = S.make_state(cfg, n, hf, seed=91)
    # adversarial: some envs exactly on cell edges / integer coordinates
    state["root_states"][:64, 0] = torch.arange(64) * 0.1 + 20.0
    state["root_states"][:64, 1] = torch.arange(64) * 0.5 + 30.0
    state["root_states"][64:128, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    oenv = O.OracleEnv(cfg, S.to_device(state, oracle_dev), hf.to(oracle_dev))
    oh = oenv._get_heights()
    px, py = oenv.last_height_indices
    env = make_env(cfg, state, hf)
    idx = env.get_height_indices().cpu()
    assert_equal(idx[..., 0], px.cpu().to(torch.int32), "px")
    assert_equal(idx[..., 1], py.cpu().to(torch.int32), "py")
    assert_equal(env._get_heights(), oh, "measured_heights")


def test_select_reset_ids_edge_cases():
    """Empty, full, ragged and unaligned flag vectors: ascending int64 ids == nonzero()."""
    import ctypes
    from isaacgymloco_b200 import _lib as L
    g = torch.Generator().manual_seed(3)
    for n, p in [(0, 0.5), (1, 1.0), (1, 0.0), (17, 0.5), (4096, 0.0), (4096, 1.0), (4096, 0.01), (16384 + 5, 0.3),
                 (65536, 0.002), (200003, 0.5)]:
        flags = (torch.rand(max(n, 1), generator=g) < p)[:n].cuda()
        ids = torch.full((max(n, 1),), -1, dtype=torch.long, device="cuda")
        cnt = torch.full((1,), -1, dtype=torch.int32, device="cuda")
        L.check(L.lib.hl_select_reset_ids(L.ptr(flags) if n else L.ptr(ids), n, L.ptr(ids), L.ptr(cnt), None, L.stream()))
        want = flags.nonzero(as_tuple=False).flatten()
        assert int(cnt.item()) == want.numel()
        assert torch.equal(ids[:want.numel()], want)


def test_full_size_properties_65536():
    """BASELINE.json's 65,536-env shard: size-independent properties of the fused step --
    ids == nonzero(reset_buf) ascending; the history really shifted; a second run from the same
    inputs is bit-identical (determinism); fused kernel == generic stage kernel."""
    from gpu_helpers import assert_close, assert_equal, make_env
    from isaacgymloco_b200 import _lib as L
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 65536
    cfg = C.aliengo("flat", num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=1234)
    noise = S.make_noise(n, seed=4)
    a = make_env(cfg, state, hf, noise=noise)
    b = make_env(cfg, state, hf, noise=noise)
    c = make_env(cfg, state, hf, noise=noise)
    old_obs = a.obs_buf.clone()
    ids, _, _ = a.post_physics_step()
    b.post_physics_step()
    assert torch.equal(ids, a.reset_buf.nonzero(as_tuple=False).flatten())
    assert torch.equal(a.obs_buf[:, 45:], old_obs[:, :225].clamp(-100, 100))
    assert torch.equal(a.obs_buf[:, :45], a.privileged_obs_buf[:, :45])
    sa, sb = a.snapshot(), b.snapshot()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    c._stages(L.ST_COUNTERS | L.ST_FRAME | L.ST_CONTACTS | L.ST_HEADING | L.ST_HEIGHTS | L.ST_TERMINATION |
              L.ST_REWARD | L.ST_OBS | L.ST_OBS_CLIP | L.ST_ROLL)
    sc = c.snapshot()
    for k in sa:
        if sa[k].dtype in (torch.bool, torch.int64):
            assert_equal(sa[k], sc[k], k)
        else:
            assert_close(sa[k], sc[k], k)


def test_philox_noise_statistics():
    """Throughput mode: in-kernel Philox noise is uniform with the reference's scale
    ((2u-1) * noise_scale_vec), independent across envs/steps, and zero where the scale is 0."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 8192
    cfg = C.aliengo("flat", num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=5)
    noisy = make_env(cfg, state, hf)                   # Philox
    cfg0 = C.aliengo("flat", num_envs=n, add_noise=False)
    clean = make_env(cfg0, state, hf)
    noisy.post_physics_step()
    clean.post_physics_step()
    d = (noisy.privileged_obs_buf - clean.privileged_obs_buf).cpu()
    nv = torch.tensor(cfg.noise_scale_vec())
    full = torch.cat([nv[:45], torch.zeros(6), nv[45:]])
    z = full == 0
    assert torch.all(d[:, z] == 0)
    u = d[:, ~z] / full[~z]                             # should be U(-1,1)
    assert abs(float(u.mean())) < 5e-3 and abs(float(u.var()) - 1 / 3) < 5e-3
    assert float(u.max()) <= 1.0 and float(u.min()) >= -1.0
    cols = torch.corrcoef(u[:, :64].T)
    assert float((cols - torch.eye(64)).abs().max()) < 0.08


def test_fixup_with_reset_zero_matches_reset_idx_bookkeeping():
    """Replay mode (bench.py): no torch reset_idx; hl_post_reset_fixup(with_reset_zero=1) applies the
    RNG-free buffer resets of LR:323-329,350,361 itself.  Two consecutive steps vs the oracle."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from oracle import torch_oracle as O
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 4096
    cfg = C.aliengo("stairs", num_envs=n)
    hf = S.make_terrain(cfg, seed=8)
    state = S.make_state(cfg, n, hf, seed=21)
    noise = S.make_noise(n, seed=22)
    oenv = O.OracleEnv(cfg, S.to_device(state, "cuda"), hf.cuda())
    env = make_env(cfg, state, hf, noise=noise)
    for step in range(2):
        oids, _, _ = oenv.post_physics_step(S.to_device(noise, "cuda"), {})   # {}: zero-only reset
        env.fused_pre_reset()
        env.fused_post_reset(with_reset_zero=True)
        env.common_step_counter += 1
        k = int(env._n_reset.item())
        assert k > 0
        assert_equal(env._reset_ids[:k], oids, "env_ids")
        compare_snapshots(env.snapshot(), oenv.snapshot())


def test_single_launch_equals_separate_kernels():
    """ids / count / terminal rows emitted by the fused kernel (decoupled look-back) are identical
    to hl_select_reset_ids + hl_terminal_rows, in Philox mode too, across several steps and for a
    ragged env count (tail CTA)."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    for n in (65536, 4096 + 37):
        cfg = C.aliengo("stairs", num_envs=n)
        hf = S.make_terrain(cfg, seed=2)
        state = S.make_state(cfg, n, hf, seed=31)
        a = make_env(cfg, state, hf)
        b = make_env(cfg, state, hf)
        b.single_launch = False
        b.refresh_buffers()
        for step in range(3):
            a.fused_pre_reset()
            b.fused_pre_reset()
            ka, kb = int(a._n_reset.item()), int(b._n_reset.item())
            assert ka == kb and ka > 0
            assert torch.equal(a._reset_ids[:ka], b._reset_ids[:kb])
            assert torch.equal(a._reset_ids[:ka], a.reset_buf.nonzero(as_tuple=False).flatten())
            assert torch.equal(a._term_priv[:ka], b._term_priv[:kb])
            assert torch.equal(a._term_amp[:ka], b._term_amp[:kb])
            # ... and to the two stand-alone kernels (hl_select_reset_ids, hl_terminal_rows)
            from isaacgymloco_b200 import _lib as L
            ids2 = torch.full((n,), -1, dtype=torch.long, device="cuda")
            cnt2 = torch.zeros(1, dtype=torch.int32, device="cuda")
            L.check(L.lib.hl_select_reset_ids(L.ptr(b.reset_buf), n, L.ptr(ids2), L.ptr(cnt2), None, L.stream()))
            assert int(cnt2.item()) == kb and torch.equal(ids2[:kb], b._reset_ids[:kb])
            rows_b = b._term_priv[:kb].clone()
            rows2 = b.compute_termination_observations(ids2[:kb])
            assert torch.equal(rows2, rows_b)
            a.fused_post_reset(with_reset_zero=True)
            b.fused_post_reset(with_reset_zero=True)
            a.common_step_counter += 1
            b.common_step_counter += 1
        sa, sb = a.snapshot(), b.snapshot()
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k


@pytest.mark.parametrize("tile", ["64", "52", "32"])
def test_every_tile_size_vs_oracle(tile, monkeypatch):
    """The fused kernel is instantiated for 64/52/32 envs per CTA (wave quantisation); force each
    and compare with the oracle, ragged env count (tail CTA) included."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    monkeypatch.setenv("HL_FUSED_EPB", tile)
    n = 3000 + 7
    cfg = C.aliengo("stairs", num_envs=n)
    hf = S.make_terrain(cfg, seed=3)
    state = S.make_state(cfg, n, hf, seed=41)
    noise = S.make_noise(n, seed=42)
    targets = S.make_reset_targets(cfg, state, hf, seed=43)
    oenv, oids, oterm, oamp = _oracle_step(cfg, state, hf, noise, targets, "cuda")
    for single in (True, False):
        env = make_env(cfg, state, hf, targets, noise)
        env.single_launch = single
        env.refresh_buffers()
        ids, term_obs, term_amp = env.post_physics_step()
        assert_equal(ids, oids, "env_ids")
        assert_close(term_obs, oterm, "termination_privileged_obs")
        assert_close(term_amp, oamp, "terminal_amp_states")
        compare_snapshots(env.snapshot(), oenv.snapshot())


def test_generic_grid_and_clipped_heights_path():
    """A non-reference scan grid (9 x 7 = 63 points) and a clip that binds on the height
    observations take the generic (any grid, clip compiled in) instantiation."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 2048
    for kw in (dict(measured_points_x=[-0.6, -0.45, -0.3, -0.15, 0., 0.15, 0.3, 0.45, 0.6],
                    measured_points_y=[-0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3]),
               dict(clip_observations=3.0)):
        cfg = C.aliengo("flat", num_envs=n, **kw)
        p = len(cfg.measured_points_x) * len(cfg.measured_points_y)
        hf = S.make_terrain(cfg, seed=4)
        state = S.make_state(cfg, n, hf, seed=51)
        noise = S.make_noise(n, seed=52, n_points=p)
        oenv, oids, oterm, oamp = _oracle_step(cfg, state, hf, noise, None, "cuda")
        env = make_env(cfg, state, hf, None, noise)
        ids, term_obs, term_amp = env.post_physics_step()
        assert env.privileged_obs_buf.shape[1] == 51 + p
        assert_equal(ids, oids, "env_ids")
        assert_close(term_obs, oterm, "termination_privileged_obs")
        compare_snapshots(env.snapshot(), oenv.snapshot())


def test_env_writes_rollout_slots_in_place():
    """bind_rollout(): the env reads its history from rollout slot `step` and writes observations /
    privileged observations into slot step+1, so record_env_step skips those two copies.  Three
    steps of a bound env + storage must equal an unbound env whose outputs are recorded by copy."""
    from gpu_helpers import make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    from isaacgymloco_b200.rollout_storage import HIMRolloutStorage
    n, t, gamma, dev = 4096, 3, 0.99, "cuda:0"
    cfg = C.aliengo("flat", num_envs=n)
    hf = S.make_terrain(cfg, seed=5)
    state = S.make_state(cfg, n, hf, seed=91)
    targets = S.make_reset_targets(cfg, state, hf, seed=92)
    envs = [make_env(cfg, state, hf, targets, S.make_noise(n, seed=93)) for _ in range(2)]
    stores = [HIMRolloutStorage(n, t, [270], [envs[0].num_privileged_obs], [12], device=dev, env_writes_slots=w) for w in (False, True)]
    envs[1].bind_rollout(stores[1])
    with pytest.raises(ValueError):
        envs[0].bind_rollout(stores[0])
    pol = [{k: v.to(dev) for k, v in S.make_transition(n, 200 + i, priv_dim=envs[0].num_privileged_obs).items()} for i in range(t)]
    for i in range(t):
        for env, st in zip(envs, stores):
            tt = HIMRolloutStorage.Transition()
            tt.observations, tt.critic_observations = env.obs_buf, env.privileged_obs_buf
            if st.env_writes_slots:
                assert tt.observations.data_ptr() == st.observations[i].data_ptr()
            else:
                tt.observations, tt.critic_observations = tt.observations.clone(), tt.critic_observations.clone()
            tt.actions, tt.values, tt.actions_log_prob = pol[i]["actions"], pol[i]["values"], pol[i]["log_prob"]
            tt.action_mean, tt.action_sigma = pol[i]["mu"], pol[i]["sigma"]
            obs, priv, rew, dones, extras, ids, term_priv = env.step(tt.actions)
            st.record_env_step(tt, rew, dones, {"time_outs": env.time_out_buf}, priv, ids, term_priv, gamma)
    for f in ("observations", "privileged_observations", "next_privileged_observations", "actions", "rewards", "dones",
              "values", "actions_log_prob", "mu", "sigma"):
        assert torch.equal(getattr(stores[0], f), getattr(stores[1], f)), f
    assert torch.equal(envs[0].obs_buf, envs[1].obs_buf) and torch.equal(envs[0].privileged_obs_buf, envs[1].privileged_obs_buf)
    assert stores[0].dones.sum() > 0
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        envs[1].step(pol[0]["actions"])
    # next rollout: slot T carries over into slot 0
    last = envs[1].obs_buf.clone()
    stores[1].clear()
    assert torch.equal(stores[1].observations[0], last)
    envs[1].step(pol[0]["actions"])
    assert envs[1].obs_buf.data_ptr() == stores[1].obs_slot(1).data_ptr()
    envs[1].bind_rollout(None)
    assert envs[1].obs_buf.data_ptr() != stores[1].obs_slot(1).data_ptr()


@pytest.mark.parametrize("oracle_dev", ["cpu", "cuda"])
def test_adversarial_thresholds(oracle_dev):
    """Inputs placed exactly on every comparison threshold of the path (SURVEY.md §8c Tier 2):
    contact-force norms at 1.0 / 0.1 / max_contact_force and one ulp above, foot Fz at 1.0,
    episode_length at 999/1000/1001, v_z at -5, base exactly on the terrain border, zero commands.
    Flags and ids must be bit-exact, floats within tolerance."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 1024
    cfg = C.aliengo("stairs", num_envs=n,
                    index_math=C.INDEX_MATH_TORCH_CPU if oracle_dev == "cpu" else C.INDEX_MATH_TORCH_CUDA)
    hf = S.make_terrain(cfg, seed=5)
    state = S.make_state(cfg, n, hf, seed=310)
    nb = cfg.num_bodies
    cf = state["contact_forces"].view(n, nb, 3)
    up = lambda x: float(np.nextafter(np.float32(x), np.float32(np.inf)))
    dn = lambda x: float(np.nextafter(np.float32(x), np.float32(-np.inf)))
    base = cfg.termination_contact_indices[0]
    pen = [b for b in cfg.penalised_contact_indices if b != base][0]
    foot = cfg.feet_indices[0]
    cf[0:64, base] = 0.0
    cf[0:16, base, 2] = 1.0                 # ||F|| == 1.0: no termination
    cf[16:32, base, 2] = up(1.0)            # one ulp above: termination
    cf[32:48, base, 0], cf[32:48, base, 1] = 0.6, 0.8      # 0.36 + 0.64 = 1.0 in exact arithmetic; fp32 decides
    cf[64:80, pen] = 0.0
    cf[64:72, pen, 1] = 0.1                 # collision threshold
    cf[72:80, pen, 1] = up(0.1)
    cf[96:112, foot] = 0.0
    cf[96:104, foot, 2] = 1.0               # contact flag threshold (> 1.0)
    cf[104:112, foot, 2] = up(1.0)
    cf[112:120, foot, 2] = float(cfg.max_contact_force)
    cf[120:128, foot, 2] = up(cfg.max_contact_force)
    state["episode_length_buf"][128:160] = torch.tensor([998, 999, 1000, 1001] * 8)
    state["root_states"][160:168, 9] = -5.0
    state["root_states"][168:176, 9] = dn(-5.0)
    state["root_states"][176:184, 0] = 0.0                  # on the lower terrain border
    state["root_states"][184:192, 1] = 0.0
    state["commands"][192:224] = 0.0
    state["root_states"][224:232, 3:7] = torch.tensor([0.0, 0.0, 1.0, 0.0])   # yaw = pi: heading wrap
    state["root_states"][232:240, 3:7] = torch.tensor([0.0, 0.0, -1.0, 0.0])
    noise = S.make_noise(n, seed=311)
    targets = S.make_reset_targets(cfg, state, hf, seed=312)
    oenv, oids, oterm, oamp = _oracle_step(cfg, state, hf, noise, targets, oracle_dev)
    env = make_env(cfg, state, hf, targets, noise)
    ids, term_obs, term_amp = env.post_physics_step()
    assert_equal(ids, oids, "env_ids")
    want = set(range(16, 32)) | {130 + 4 * k for k in range(8)} | {131 + 4 * k for k in range(8)} | set(range(168, 176))
    assert want <= set(ids.cpu().tolist()), "threshold envs that must reset"
    assert_close(term_obs, oterm, "termination_privileged_obs")
    assert_close(term_amp, oamp, "terminal_amp_states")
    compare_snapshots(env.snapshot(), oenv.snapshot())
