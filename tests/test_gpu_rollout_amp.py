"""-m gpu: GAE/returns and the AMP data path through the C ABI vs the reference fixtures and
the torch oracle.  returns: bit-exact (same op order, no contraction); advantages: rel 1e-5;
AMP frame indices / lerp columns / pair gathers / disc input: bit-exact; slerp columns rel 1e-5."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _storage(n, t, r):
    from isaacgymloco_b200.rollout_storage import HIMRolloutStorage
    st = HIMRolloutStorage(n, t, [270], [238], [12], device="cuda:0")
    st.rewards.copy_(r["rewards"])
    st.values.copy_(r["values"])
    st.dones.copy_(r["dones"])
    return st


@pytest.mark.parametrize("name", ["a", "b", "c", "alldone"])
def test_gae_vs_reference_golden(name):
    from isaacgymloco_b200 import synthetic as S
    gold = load_golden("gae.npz")
    n, t, seed, gamma, lam = gold[f"{name}_meta"]
    n, t = int(n), int(t)
    r = S.make_rollout(n, t, int(seed))
    if name == "alldone":
        r["dones"][:] = 1
    st = _storage(n, t, r)
    st.compute_returns(r["last_values"].cuda(), gamma, lam)
    assert st.returns.shape == (t, n, 1) and st.advantages.shape == (t, n, 1)
    np.testing.assert_array_equal(st.returns.cpu().numpy(), gold[f"{name}_returns"])
    np.testing.assert_allclose(st.advantages.cpu().numpy(), gold[f"{name}_advantages"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("n,t", [(4096, 24), (4096, 100), (65536, 24)])
def test_gae_vs_oracle_full_size(n, t):
    """configs[0]/[1] (4096 x 24), the real rollout length (T=100) and the 65,536-env shard."""
    from isaacgymloco_b200 import synthetic as S
    from oracle import torch_oracle as O
    r = S.make_rollout(n, t, 7)
    ret, adv = O.compute_returns(r["rewards"], r["values"], r["dones"], r["last_values"], 0.99, 0.95)
    st = _storage(n, t, r)
    st.compute_returns(r["last_values"].cuda(), 0.99, 0.95)
    assert torch.equal(st.returns.cpu(), ret)
    np.testing.assert_allclose(st.advantages.cpu().numpy(), adv.numpy(), rtol=1e-5, atol=2e-6)
    # size-independent properties: normalised advantages have mean 0 / unbiased std 1, and
    # returns - values recovers the un-normalised advantage up to that affine map
    a = st.advantages.double()
    assert abs(float(a.mean())) < 1e-5 and abs(float(a.std()) - 1.0) < 1e-5
    raw = (st.returns - st.values).double()
    np.testing.assert_allclose(((raw - raw.mean()) / (raw.std() + 1e-8)).cpu().numpy(), a.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_gae_linearity_and_terminal_cut():
    """Domain properties: with dones=0, GAE is linear in (rewards, values); a done at step t makes
    returns[<=t] independent of everything after t."""
    from isaacgymloco_b200 import synthetic as S
    n, t = 1024, 16
    r = S.make_rollout(n, t, 11)
    r["dones"][:] = 0
    st = _storage(n, t, r)
    st.compute_returns(r["last_values"].cuda(), 0.99, 0.95)
    base = st.returns.clone()
    r2 = {k: (v * 2 if v.dtype.is_floating_point else v) for k, v in r.items()}
    st2 = _storage(n, t, r2)
    st2.compute_returns(r2["last_values"].cuda(), 0.99, 0.95)
    np.testing.assert_allclose(st2.returns.cpu().numpy(), 2 * base.cpu().numpy(), rtol=1e-6, atol=1e-6)
    r3 = {k: v.clone() for k, v in r.items()}
    r3["dones"][7] = 1
    st3 = _storage(n, t, r3)
    st3.compute_returns(r3["last_values"].cuda(), 0.99, 0.95)
    r4 = {k: v.clone() for k, v in r3.items()}
    r4["rewards"][8:] += 5.0
    r4["values"][8:] -= 3.0
    r4["last_values"] += 1.0
    st4 = _storage(n, t, r4)
    st4.compute_returns(r4["last_values"].cuda(), 0.99, 0.95)
    assert torch.equal(st3.returns[:8], st4.returns[:8])


def _loader(gold, preload=False, n_pre=0):
    from isaacgymloco_b200.motion_loader import AMPLoader
    k = len(gold["frame_durations"])
    tables = dict(frames=[gold[f"clip{i}"] for i in range(k)], frame_durations=gold["frame_durations"],
                  weights=gold["weights_raw"], names=[str(x) for x in gold["clip_names"]])
    return AMPLoader("cuda:0", 0.02, preload_transitions=preload, num_preload_transitions=n_pre, clip_tables=tables)


# ----------------------------------------------------------------------------- rollout-step recording (§8f rank 1)
def _transition(tr, dev="cuda:0"):
    from isaacgymloco_b200.rollout_storage import HIMRolloutStorage
    t = HIMRolloutStorage.Transition()
    t.observations, t.critic_observations = tr["obs"].to(dev), tr["critic_obs"].to(dev)
    t.actions, t.values, t.actions_log_prob = tr["actions"].to(dev), tr["values"].to(dev), tr["log_prob"].to(dev)
    t.action_mean, t.action_sigma = tr["mu"].to(dev), tr["sigma"].to(dev)
    return t


def _record(st, tr, gamma, dev="cuda:0", **kw):
    st.record_env_step(_transition(tr, dev), tr["rewards"].to(dev), tr["dones"].to(dev), {"time_outs": tr["time_outs"].to(dev)},
                       tr["privileged_obs"].to(dev), tr["termination_ids"].to(dev),
                       tr["termination_privileged_obs"].to(dev), gamma, **kw)


@pytest.mark.parametrize("name", ["a", "b", "one"])
def test_record_env_step_vs_reference_golden(name):
    """hl_record_transition vs the reference's HIMPPO.process_env_step + add_transitions: bit-exact."""
    from isaacgymloco_b200.rollout_storage import HIMRolloutStorage
    from test_oracle_golden import RECORD_FIELDS, record_case_steps
    gold = load_golden("record.npz")
    n, t, seed, gamma = gold[f"{name}_meta"]
    n, t = int(n), int(t)
    st = HIMRolloutStorage(n, t, [270], [238], [12], device="cuda:0")
    for tr in record_case_steps(name, n, t, int(seed)):
        _record(st, tr, gamma)
    assert st.step == t
    for f in RECORD_FIELDS:
        np.testing.assert_array_equal(getattr(st, f).cpu().numpy(), gold[f"{name}_{f}"], err_msg=f)
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        _record(st, record_case_steps(name, n, 1, int(seed))[0], gamma)


def test_record_env_step_full_size_and_variants():
    """65,536 envs vs the oracle run on the same device; add_transitions drop-in (pre-patched
    transition); over-allocated id buffers with a device-side count; unsorted ids; in-place slots."""
    import oracle.torch_oracle as O
    from isaacgymloco_b200 import synthetic as S
    from isaacgymloco_b200.rollout_storage import HIMRolloutStorage
    from test_oracle_golden import RECORD_FIELDS
    n, t, gamma, dev = 65536, 2, 0.99, "cuda:0"
    steps = [{k: v.to(dev) for k, v in S.make_transition(n, 70 + i).items()} for i in range(t)]
    ref = {f: torch.zeros(t, n, {"observations": 270, "privileged_observations": 238, "next_privileged_observations": 238,
                                  "actions": 12, "mu": 12, "sigma": 12}.get(f, 1), device=dev,
                          dtype=torch.uint8 if f == "dones" else torch.float32) for f in RECORD_FIELDS}
    for i, tr in enumerate(steps):
        O.record_env_step(ref, i, tr, gamma)

    def check(st, msg):
        for f in RECORD_FIELDS:
            assert torch.equal(getattr(st, f), ref[f]), f"{msg}: {f}"

    # fused call
    st = HIMRolloutStorage(n, t, [270], [238], [12], device=dev)
    for tr in steps:
        _record(st, tr, gamma)
    check(st, "fused")
    # drop-in add_transitions: the caller did the patch and the bootstrap (reference flow)
    st = HIMRolloutStorage(n, t, [270], [238], [12], device=dev)
    for tr in steps:
        tt = _transition(tr)
        nxt = tr["privileged_obs"].clone()
        nxt[tr["termination_ids"]] = tr["termination_privileged_obs"]
        tt.next_critic_observations = nxt
        tt.rewards = tr["rewards"] + gamma * torch.squeeze(tr["values"] * tr["time_outs"].unsqueeze(1), 1)
        tt.dones = tr["dones"]
        st.add_transitions(tt)
    check(st, "add_transitions")
    # over-allocated id / row buffers + device-side count (what a sync-free env.step hands over)
    st = HIMRolloutStorage(n, t, [270], [238], [12], device=dev)
    for tr in steps:
        k = tr["termination_ids"].numel()
        ids = torch.full((n,), -1, dtype=torch.int64, device=dev)
        ids[:k] = tr["termination_ids"]
        rows = torch.full((k + 100, 238), float("nan"), device=dev)
        rows[:k] = tr["termination_privileged_obs"]
        st.record_env_step(_transition(tr), tr["rewards"], tr["dones"], {"time_outs": tr["time_outs"]}, tr["privileged_obs"],
                           ids, rows, gamma, termination_count=torch.tensor([k], dtype=torch.int32, device=dev))
    check(st, "device count")
    # unsorted ids
    st = HIMRolloutStorage(n, t, [270], [238], [12], device=dev)
    for tr in steps:
        perm = torch.randperm(tr["termination_ids"].numel(), device=dev)
        tr2 = dict(tr, termination_ids=tr["termination_ids"][perm], termination_privileged_obs=tr["termination_privileged_obs"][perm])
        _record(st, tr2, gamma, assume_sorted=False)
    check(st, "unsorted")
    # in-place: the env wrote obs / critic obs straight into the slot -> those copies are skipped
    st = HIMRolloutStorage(n, t, [270], [238], [12], device=dev)
    for i, tr in enumerate(steps):
        st.observations[i].copy_(tr["obs"])
        st.privileged_observations[i].copy_(tr["critic_obs"])
        tt = _transition(tr)
        tt.observations, tt.critic_observations = st.observations[i], st.privileged_observations[i]
        st.record_env_step(tt, tr["rewards"], tr["dones"], {"time_outs": tr["time_outs"]}, tr["privileged_obs"],
                           tr["termination_ids"], tr["termination_privileged_obs"], gamma)
    check(st, "in place")
    # no resets, no time-out info, odd env count (ragged last tile)
    n2 = 1000 + 7
    tr = {k: v.to(dev) for k, v in S.make_transition(n2, 5, reset_frac=0.0).items()}
    st = HIMRolloutStorage(n2, 1, [270], [238], [12], device=dev)
    st.record_env_step(_transition(tr), tr["rewards"], tr["dones"], {}, tr["privileged_obs"], tr["termination_ids"],
                       tr["termination_privileged_obs"], gamma)
    assert torch.equal(st.next_privileged_observations[0], tr["privileged_obs"])
    assert torch.equal(st.rewards[0, :, 0], tr["rewards"]) and torch.equal(st.observations[0], tr["obs"])


# ----------------------------------------------------------------------------- fused minibatch gather (§8f rank 3)
def _filled(n, t, seed, dev="cuda:0"):
    from isaacgymloco_b200 import synthetic as S
    from isaacgymloco_b200.rollout_storage import HIMRolloutStorage
    st = HIMRolloutStorage(n, t, [270], [238], [12], device=dev)
    src = S.make_filled_storage(n, t, seed)
    for k, v in src.items():
        getattr(st, k).copy_(v)
    return st, src


def test_mini_batch_generator_vs_reference_golden():
    """hl_minibatch_gather behind mini_batch_generator vs the reference generator: bit-exact rows,
    same yield order and shapes."""
    gold = load_golden("minibatch.npz")
    n, t, nmb, epochs, seed = (int(x) for x in gold["meta"])
    st, _ = _filled(n, t, seed)
    batches = list(st.mini_batch_generator(nmb, epochs, indices=torch.from_numpy(gold["indices"]).cuda()))
    assert len(batches) == nmb * epochs
    for bi, b in enumerate(batches):
        assert len(b) == 10
        for fi, x in enumerate(b):
            np.testing.assert_array_equal(x.cpu().numpy(), gold[f"b{bi}_f{fi}"], err_msg=f"batch {bi} field {fi}")


def test_mini_batch_gather_full_size_properties():
    """16,384 envs x 24 steps, 4 minibatches: every fused gather equals torch indexing; the
    default (randperm) path yields each row exactly once per epoch; ragged / empty index lists."""
    from oracle import torch_oracle as O
    n, t, nmb = 16384, 24, 4
    st, _ = _filled(n, t, 3)
    g = torch.Generator(device="cuda").manual_seed(5)
    perm = torch.randperm(n * t, device="cuda", generator=g)
    ref = {k: getattr(st, k) for k in O.MINIBATCH_ORDER}
    for b, rb in zip(st.mini_batch_generator(nmb, 1, indices=perm), O.mini_batches(ref, nmb, 1, perm)):
        for x, y in zip(b, rb):
            assert x.shape == y.shape and torch.equal(x, y)
    # default path: a permutation of all rows (checksum of the values column)
    tot = torch.zeros((), dtype=torch.float64, device="cuda")
    rows = 0
    for b in st.mini_batch_generator(nmb, 1):
        tot += b[4].double().sum()
        rows += b[4].shape[0]
    assert rows == n * t
    assert abs(float(tot) - float(st.values.double().sum())) < 1e-6 * n * t
    # ragged (not a multiple of the 32-row tile), duplicate and empty index lists
    idx = torch.tensor([5, 5, n * t - 1, 0, 17] * 13, device="cuda")
    out = st.gather_batch(idx)
    for x, k in zip(out, O.MINIBATCH_ORDER):
        assert torch.equal(x, getattr(st, k).flatten(0, 1)[idx])
    assert st.gather_batch(idx[:0])[0].shape == (0, 270)


def test_replay_buffer_vs_reference_golden():
    """ReplayBuffer (hl_ring_insert + hl_minibatch_gather) vs the reference's ReplayBuffer: inserts
    that wrap and that exceed the ring, step / num_samples bookkeeping, sampled minibatches."""
    from isaacgymloco_b200.replay_buffer import ReplayBuffer
    from test_oracle_golden import replay_inputs
    gold = load_golden("replay.npz")
    rb = ReplayBuffer(30, 16, "cuda:0")
    for i, (a, b) in enumerate(replay_inputs()):
        rb.insert(a.cuda(), b.cuda())
        np.testing.assert_array_equal(rb.states.cpu().numpy(), gold[f"states_{i}"])
        np.testing.assert_array_equal(rb.next_states.cpu().numpy(), gold[f"next_{i}"])
        assert [rb.step, rb.num_samples] == list(gold[f"meta_{i}"])
    np.random.seed(123)
    for j, (s_, n_) in enumerate(rb.feed_forward_generator(3, 7)):
        np.testing.assert_array_equal(s_.cpu().numpy(), gold[f"mb_s{j}"])
        np.testing.assert_array_equal(n_.cpu().numpy(), gold[f"mb_n{j}"])
    with pytest.raises(RuntimeError):
        rb.insert(torch.zeros(40, 30, device="cuda"), torch.zeros(40, 30, device="cuda"))
    # config-4 sizes against the oracle on the same device
    from oracle import torch_oracle as O
    big, ob = ReplayBuffer(30, 100000, "cuda:0"), O.OracleReplayBuffer(30, 100000, "cuda")
    g = torch.Generator(device="cuda").manual_seed(4)
    for _ in range(8):
        a, b = torch.randn(16384, 30, device="cuda", generator=g), torch.randn(16384, 30, device="cuda", generator=g)
        big.insert(a, b); ob.insert(a, b)
    assert torch.equal(big.states, ob.states) and torch.equal(big.next_states, ob.next_states)
    assert (big.step, big.num_samples) == (ob.step, ob.num_samples)
    np.random.seed(9); x = list(big.feed_forward_generator(2, 4096))
    np.random.seed(9); y = list(ob.feed_forward_generator(2, 4096))
    for (s1, n1), (s2, n2) in zip(x, y):
        assert torch.equal(s1, s2) and torch.equal(n1, n2)


def test_amp_frame_blend_vs_reference_golden():
    gold = load_golden("amp.npz")
    ld = _loader(gold)
    assert ld.observation_dim == 30 and ld.num_motions == 7
    np.testing.assert_array_equal(ld.trajectory_lens, gold["lens"])
    np.testing.assert_array_equal(ld.trajectory_weights, gold["weights"])
    out, lo, hi = ld.get_full_frame_at_time_batch(gold["blend_idx"], gold["blend_times"], return_indices=True)
    p = gold["blend_times"] / gold["lens"][gold["blend_idx"]]
    pn = p * gold["num_frames"][gold["blend_idx"]]
    np.testing.assert_array_equal(lo.cpu().numpy(), np.floor(pn).astype(np.int32))
    np.testing.assert_array_equal(hi.cpu().numpy(), np.ceil(pn).astype(np.int32))
    got, want = out.cpu().numpy(), gold["blend_frames"]
    lerp_cols = [c for c in range(49) if not 3 <= c < 7]
    np.testing.assert_array_equal(got[:, lerp_cols], want[:, lerp_cols])
    np.testing.assert_allclose(got[:, 3:7], want[:, 3:7], rtol=1e-5, atol=1e-6, equal_nan=True)


def test_amp_loader_from_binary_cache(tmp_path):
    """save_cache -> from_cache gives the same tables and the same blended frames."""
    gold = load_golden("amp.npz")
    a = _loader(gold)
    path = str(tmp_path / "clips.npz")
    a.save_cache(path)
    from isaacgymloco_b200.motion_loader import AMPLoader
    b = AMPLoader.from_cache(path, "cuda:0", 0.02)
    assert b.trajectory_names == a.trajectory_names
    np.testing.assert_array_equal(b.trajectory_weights, a.trajectory_weights)
    np.testing.assert_array_equal(b.trajectory_lens, a.trajectory_lens)
    assert torch.equal(a.all_trajectories_full, b.all_trajectories_full)
    fa = a.get_full_frame_at_time_batch(gold["blend_idx"], gold["blend_times"])
    fb = b.get_full_frame_at_time_batch(gold["blend_idx"], gold["blend_times"])
    assert torch.equal(fa, fb)


def test_amp_frame_blend_config4_size():
    """configs[3]: 16,384 samples per batch (and a 2e5-sample preload) vs the oracle."""
    from oracle import torch_oracle as O
    gold = load_golden("amp.npz")
    ld = _loader(gold)
    tab = O.OracleMotionTable([torch.from_numpy(gold[f"clip{i}"]) for i in range(7)], gold["frame_durations"],
                              gold["weights_raw"], 0.02)
    np.random.seed(5)
    for b in (16384, 200000):
        idx = ld.weighted_traj_idx_sample_batch(b)
        times = ld.traj_time_sample_batch(idx)
        want, lo, hi = tab.get_full_frame_at_time_batch(idx, times)
        got, glo, ghi = ld.get_full_frame_at_time_batch(idx, times, return_indices=True)
        np.testing.assert_array_equal(glo.cpu().numpy(), lo.astype(np.int32))
        np.testing.assert_array_equal(ghi.cpu().numpy(), hi.astype(np.int32))
        lerp_cols = [c for c in range(49) if not 3 <= c < 7]
        np.testing.assert_array_equal(got.cpu().numpy()[:, lerp_cols], want.numpy()[:, lerp_cols])
        np.testing.assert_allclose(got.cpu().numpy()[:, 3:7], want.numpy()[:, 3:7], rtol=1e-5, atol=1e-6, equal_nan=True)


def test_amp_pairs_and_disc_reward_vs_reference_golden():
    from isaacgymloco_b200.amp_discriminator import AMPDiscriminator, Normalizer
    gold = load_golden("amp.npz")
    ld = _loader(gold)
    ld.preload_transitions = True
    ld.preloaded_s = torch.from_numpy(gold["pre_s"]).cuda()
    ld.preloaded_s_next = torch.from_numpy(gold["pre_s_next"]).cuda()
    for k in (0, 1):
        s, sn = ld.gather_pairs(gold[f"pair_idx{k}"])
        np.testing.assert_array_equal(s.cpu().numpy(), gold[f"pair_s{k}"])
        np.testing.assert_array_equal(sn.cpu().numpy(), gold[f"pair_sn{k}"])
    # the generator consumes np.random exactly like the reference's (same idx stream)
    np.random.seed(33)
    pairs = list(ld.feed_forward_generator(2, 512))
    np.testing.assert_array_equal(pairs[0][0].cpu().numpy(), gold["pair_s0"])
    np.testing.assert_array_equal(pairs[1][1].cpu().numpy(), gold["pair_sn1"])
    # normaliser: float64 device moments vs the reference's numpy update on fp32 batches
    # (the reference accumulates np.mean/np.var in fp32 => agreement to ~1e-6 rel, not bit-exact)
    norm = Normalizer(30)
    norm.update(pairs[0][0])
    norm.update(pairs[1][0])
    np.testing.assert_allclose(norm.mean.cpu().numpy(), gold["norm_mean"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(norm.var.cpu().numpy(), gold["norm_var"], rtol=2e-5, atol=1e-6)
    assert abs(norm.count - float(gold["norm_count"])) < 1e-9
    # discriminator reward with the golden normaliser state
    norm.mean = torch.from_numpy(gold["norm_mean"]).cuda()
    norm.var = torch.from_numpy(gold["norm_var"]).cuda()
    torch.manual_seed(34)
    disc = AMPDiscriminator(60, 0.01, [64, 32], "cuda:0", task_reward_lerp=0.3)
    sd = {k: torch.from_numpy(gold["disc_" + k.replace(".", "_")]) for k in disc.state_dict()}
    disc.load_state_dict(sd)
    s, sn = torch.from_numpy(gold["disc_s"]).cuda(), torch.from_numpy(gold["disc_sn"]).cuda()
    x = disc.assemble_input(s, sn, norm)
    np.testing.assert_array_equal(x.cpu().numpy(), gold["disc_x"])
    r, d = disc.predict_amp_reward(s, sn, torch.from_numpy(gold["disc_task_r"]).cuda(), normalizer=norm)
    assert r.shape == (256,) and d.shape == (256, 1)
    np.testing.assert_allclose(d.cpu().numpy(), gold["disc_d"], rtol=1e-4, atol=1e-5)   # cuBLAS vs CPU GEMM
    # epilogue alone, on the golden logits: exact op order => bit-exact
    from isaacgymloco_b200 import _lib as L
    rr = torch.empty(256, device="cuda")
    dd = torch.from_numpy(gold["disc_d"]).cuda().contiguous()
    tr = torch.from_numpy(gold["disc_task_r"]).cuda()
    L.check(L.lib.hl_amp_reward(L.ptr(dd), L.ptr(tr), 0.01, 0.3, L.ptr(rr), 256, L.stream()))
    np.testing.assert_array_equal(rr.cpu().numpy(), gold["disc_r"])


def test_amp_terminal_patch_config4():
    """configs[3]: discriminator-input batch for 16,384 envs with the runner's terminal patch."""
    from isaacgymloco_b200.amp_discriminator import AMPDiscriminator, Normalizer
    from oracle import torch_oracle as O
    n = 16384
    g = torch.Generator().manual_seed(9)
    s, sn = torch.randn(n, 30, generator=g), torch.randn(n, 30, generator=g)
    ids = (torch.rand(n, generator=g) < 0.01).nonzero().flatten()
    term = torch.randn(len(ids), 30, generator=g)
    mean, var = np.random.default_rng(1).normal(size=30), np.random.default_rng(2).uniform(0.5, 2, size=30)
    want_next = sn.clone()
    want_next[ids] = term
    want = O.amp_disc_input(s, want_next, mean, var)
    norm = Normalizer(30)
    norm.mean, norm.var = torch.from_numpy(mean).cuda(), torch.from_numpy(var).cuda()
    disc = AMPDiscriminator(60, 0.01, [1024, 512], "cuda:0", task_reward_lerp=0.3)
    x, patched = disc.assemble_input(s.cuda(), sn.cuda(), norm, ids.cuda(), term.cuda(), return_patched=True)
    np.testing.assert_array_equal(patched.cpu().numpy(), want_next.numpy())
    np.testing.assert_array_equal(x.cpu().numpy(), want.numpy())
    x0 = disc.assemble_input(s.cuda(), sn.cuda(), None)
    np.testing.assert_array_equal(x0.cpu().numpy(), torch.cat([s, sn], -1).numpy())


def test_amp_blend_adversarial_times():
    """Times exactly on frame boundaries (blend 0), at t=0 and at the clip end (lo == hi == last
    frame), plus antipodal / identical quaternion pairs through the slerp (its 1/angle quirk included)."""
    from oracle import torch_oracle as O
    gold = load_golden("amp.npz")
    ld = _loader(gold)
    tab = O.OracleMotionTable([torch.from_numpy(gold[f"clip{i}"]) for i in range(7)], gold["frame_durations"],
                              gold["weights_raw"], 0.02)
    idx, times = [], []
    for c in range(7):
        length, nf = tab.trajectory_lens[c], int(tab.trajectory_num_frames[c])
        for k in (0, 1, nf // 2, nf - 2, nf - 1):
            idx.append(c); times.append(length * k / nf)                 # p*n lands on (or an ulp off) an integer
        idx += [c, c]
        times += [0.0, np.nextafter(length * (nf - 1) / nf, 0.0)]
    idx, times = np.array(idx), np.array(times, dtype=np.float64)
    # k = nf-1 can land one ulp above the last frame (then the reference itself raises): keep the valid ones
    ok = np.ceil(times / tab.trajectory_lens[idx] * tab.trajectory_num_frames[idx]) <= tab.trajectory_num_frames[idx] - 1
    assert ok.sum() >= len(ok) - 7
    idx, times = idx[ok], times[ok]
    with pytest.raises(IndexError):    # the reference indexes past the clip for t = len (ML:243-244)
        ld.get_full_frame_at_time_batch(np.array([0]), np.array([tab.trajectory_lens[0]]))
    want, lo, hi = tab.get_full_frame_at_time_batch(idx, times)
    got, glo, ghi = ld.get_full_frame_at_time_batch(idx, times, return_indices=True)
    np.testing.assert_array_equal(glo.cpu().numpy(), lo.astype(np.int32))
    np.testing.assert_array_equal(ghi.cpu().numpy(), hi.astype(np.int32))
    lerp_cols = [c for c in range(49) if not 3 <= c < 7]
    np.testing.assert_array_equal(got.cpu().numpy()[:, lerp_cols], want.numpy()[:, lerp_cols])
    np.testing.assert_allclose(got.cpu().numpy()[:, 3:7], want.numpy()[:, 3:7], rtol=1e-5, atol=1e-6, equal_nan=True)
