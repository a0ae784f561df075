"""CPU: the Tier-1 oracle (oracle/torch_oracle.py) against fixtures minted from the real
reference (oracle/make_goldens.py).  Flags, ids and height-field indices are compared bit-exact;
floats at fp32 rounding level (the oracle repeats the reference's op sequence on the same
device, so most are bit-equal too)."""
import numpy as np
import pytest
import torch

from conftest import build_case_inputs, load_golden
from oracle import torch_oracle as O
from oracle.make_goldens import ENV_CASES

EXACT = ("reset_buf", "time_out_buf", "contact_filt", "last_contacts", "episode_length_buf")
FLOATS = ("base_lin_vel", "base_ang_vel", "projected_gravity", "measured_heights", "rew_buf",
          "feet_air_time", "commands", "obs_buf", "privileged_obs_buf", "last_actions",
          "last_last_actions", "last_dof_pos", "last_dof_vel", "last_torques", "last_root_vel",
          "episode_sums")


@pytest.mark.parametrize("case", ENV_CASES, ids=[c["name"] for c in ENV_CASES])
def test_env_case(case):
    gold = load_golden(f"env_{case['name']}.npz")
    cfg, hf, state, noise, targets, delayed, chk = build_case_inputs(case)
    assert chk == gold["input_checksum"], "synthetic generator drifted from the minted inputs"
    env = O.OracleEnv(cfg, state, hf)
    tq = torch.stack([env._compute_torques(delayed[:, k]) for k in range(4)], dim=1)
    np.testing.assert_allclose(tq.numpy(), gold["torques4"], rtol=1e-6, atol=1e-6)
    if not cfg.is_plane:
        env._get_heights()
        px, py = env.last_height_indices
        np.testing.assert_array_equal(px.numpy(), gold["px"].astype(np.int64))
        np.testing.assert_array_equal(py.numpy(), gold["py"].astype(np.int64))
    ids, term_obs, term_amp = env.post_physics_step(noise, None if case.get("no_reset") else targets)
    np.testing.assert_array_equal(ids.numpy(), gold["env_ids"])
    np.testing.assert_allclose(term_obs.numpy(), gold["term_obs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(term_amp.numpy(), gold["term_amp"], rtol=1e-6, atol=1e-6)
    snap = env.snapshot()
    for k in EXACT:
        np.testing.assert_array_equal(snap[k].numpy(), gold[k], err_msg=k)
    for k in FLOATS:
        np.testing.assert_allclose(snap[k].numpy(), gold[k], rtol=1e-6, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("name", ["a", "b", "c", "alldone"])
def test_gae(name):
    from isaacgymloco_b200 import synthetic as S
    gold = load_golden("gae.npz")
    n, t, seed, gamma, lam = gold[f"{name}_meta"]
    r = S.make_rollout(int(n), int(t), int(seed))
    if name == "alldone":
        r["dones"][:] = 1
    ret, adv = O.compute_returns(r["rewards"], r["values"], r["dones"], r["last_values"], gamma, lam)
    np.testing.assert_array_equal(ret.numpy(), gold[f"{name}_returns"])
    np.testing.assert_array_equal(adv.numpy(), gold[f"{name}_advantages"])


RECORD_FIELDS = ("observations", "privileged_observations", "next_privileged_observations", "actions", "rewards", "dones",
                 "values", "actions_log_prob", "mu", "sigma")


def record_case_steps(name, n, t, seed):
    """The synthetic per-step inputs oracle/make_goldens.py::mint_record fed the reference."""
    from isaacgymloco_b200 import synthetic as S
    return [S.make_transition(n, seed * 10 + step, reset_frac=(1.0 if name == "one" and step == 1 else 0.05))
            for step in range(t)]


@pytest.mark.parametrize("name", ["a", "b", "one"])
def test_record_env_step(name):
    """Oracle restatement of runner patch + process_env_step + add_transitions vs the reference's
    own HIMPPO / HIMRolloutStorage (bit-exact: copies plus one bootstrap expression)."""
    gold = load_golden("record.npz")
    n, t, seed, gamma = gold[f"{name}_meta"]
    n, t = int(n), int(t)
    st = {f: torch.zeros(gold[f"{name}_{f}"].shape, dtype=torch.uint8 if f == "dones" else torch.float32)
          for f in RECORD_FIELDS}
    for step, tr in enumerate(record_case_steps(name, n, t, int(seed))):
        O.record_env_step(st, step, tr, gamma)
    for f in RECORD_FIELDS:
        np.testing.assert_array_equal(st[f].numpy(), gold[f"{name}_{f}"], err_msg=f)


def test_mini_batch_generator():
    """Oracle restatement of mini_batch_generator vs the reference's own generator (same permutation)."""
    from isaacgymloco_b200 import synthetic as S
    gold = load_golden("minibatch.npz")
    n, t, nmb, epochs, seed = (int(x) for x in gold["meta"])
    st = S.make_filled_storage(n, t, seed)
    batches = list(O.mini_batches(st, nmb, epochs, torch.from_numpy(gold["indices"])))
    assert len(batches) == nmb * epochs
    for bi, b in enumerate(batches):
        for fi, x in enumerate(b):
            np.testing.assert_array_equal(x.numpy(), gold[f"b{bi}_f{fi}"], err_msg=f"batch {bi} field {fi}")


REPLAY_INSERTS = (5, 9, 4, 12, 1, 17, 3)   # mirrors oracle/make_goldens.py


def replay_inputs(obs_dim=30, seed=88):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(k, obs_dim, generator=g), torch.randn(k, obs_dim, generator=g)) for k in REPLAY_INSERTS]


def test_replay_buffer():
    gold = load_golden("replay.npz")
    rb = O.OracleReplayBuffer(30, 16)
    for i, (a, b) in enumerate(replay_inputs()):
        rb.insert(a, b)
        np.testing.assert_array_equal(rb.states.numpy(), gold[f"states_{i}"])
        np.testing.assert_array_equal(rb.next_states.numpy(), gold[f"next_{i}"])
        assert [rb.step, rb.num_samples] == list(gold[f"meta_{i}"])
    np.random.seed(123)
    for j, (s_, n_) in enumerate(rb.feed_forward_generator(3, 7)):
        np.testing.assert_array_equal(s_.numpy(), gold[f"mb_s{j}"])
        np.testing.assert_array_equal(n_.numpy(), gold[f"mb_n{j}"])


def _table(gold):
    clips = [torch.from_numpy(gold[f"clip{i}"]) for i in range(len(gold["frame_durations"]))]
    return O.OracleMotionTable(clips, gold["frame_durations"], gold["weights_raw"], 0.02)


def test_amp_blend():
    gold = load_golden("amp.npz")
    tab = _table(gold)
    np.testing.assert_array_equal(tab.trajectory_lens, gold["lens"])
    np.testing.assert_array_equal(tab.trajectory_weights, gold["weights"])
    frames, lo, hi = tab.get_full_frame_at_time_batch(gold["blend_idx"], gold["blend_times"])
    np.testing.assert_array_equal(frames.numpy(), gold["blend_frames"])


def test_amp_pairs_and_reward():
    gold = load_golden("amp.npz")
    pre_s, pre_sn = torch.from_numpy(gold["pre_s"]), torch.from_numpy(gold["pre_s_next"])
    for k in (0, 1):
        s, sn = O.amp_pairs(pre_s, pre_sn, gold[f"pair_idx{k}"])
        np.testing.assert_array_equal(s.numpy(), gold[f"pair_s{k}"])
        np.testing.assert_array_equal(sn.numpy(), gold[f"pair_sn{k}"])
    # normaliser moments (utils.py:90-110)
    mean, var, count = np.zeros(30), np.ones(30), 1e-4
    for k in (0, 1):
        mean, var, count = O.running_moments_update(mean, var, count, gold[f"pair_s{k}"])
    np.testing.assert_array_equal(mean, gold["norm_mean"])
    np.testing.assert_array_equal(var, gold["norm_var"])
    x = O.amp_disc_input(torch.from_numpy(gold["disc_s"]), torch.from_numpy(gold["disc_sn"]),
                         gold["norm_mean"], gold["norm_var"])
    np.testing.assert_array_equal(x.numpy(), gold["disc_x"])
    r = O.amp_reward_from_logit(torch.from_numpy(gold["disc_d"]), torch.from_numpy(gold["disc_task_r"]),
                                0.01, 0.3)
    np.testing.assert_array_equal(r.numpy(), gold["disc_r"])


@pytest.mark.parametrize("case", __import__("oracle.make_goldens", fromlist=["RESET_CASES"]).RESET_CASES, ids=lambda c: c["name"])
def test_oracle_reset_idx_reproduces_the_reference(case):
    """reset_idx with its draws (LR:288-361): the fixtures come from the reference's own reset_idx with
    torch_rand_float / rand_like / randint_like fed from a (N, 44) uniform table in call order; the restatement
    consuming the same table by column must give the same rows."""
    from oracle.make_goldens import reset_inputs, input_checksum
    from oracle import torch_oracle as O
    gold = load_golden(f"reset_{case['name']}.npz")
    cfg, hf, state, u, ter, ids = reset_inputs(case)
    # (sums of doubles: the summation order depends on the thread count, hence isclose)
    assert np.isclose(input_checksum(state), gold["input_checksum"], rtol=1e-12) and np.isclose(float(u.double().sum()), float(gold["u_checksum"]), rtol=1e-12)
    np.testing.assert_array_equal(ids.numpy(), gold["ids"])
    env = O.OracleEnv(cfg, state, hf)
    env.env_origins = ter["env_origins"].clone()
    terrain = dict(origins=ter["origins"], types=ter["types"], max_level=cfg.num_rows, env_length=cfg.terrain_length,
                   max_episode_length_s=cfg.episode_length_s)
    env.reset_idx_draw(ids, u, custom_origins=True, terrain=terrain)
    np.testing.assert_array_equal(env.terrain_levels.numpy(), gold["terrain_levels"])
    for k in ("root_states", "dof_state", "commands", "Kp_factors", "Kd_factors", "motor_strength_factors", "env_origins"):
        np.testing.assert_allclose(getattr(env, k).numpy().reshape(gold[k].shape), gold[k], rtol=1e-6, atol=1e-7, err_msg=k)
