"""-m gpu: the persistent, role-pipelined form of the fused step (HL_FUSED_IMPL=persist,
isaacgymloco_b200/csrc/hl_persist_kernel.inc: TMA bulk loads, mbarrier hand-offs, ticket-ordered tiles)
against the tiled default kernel on the same inputs, and against the torch oracle.

Bit-exact between the two forms: ids, terminal rows, flags, counters, heights, observations, the last_* roll.
`rew_buf` / `episode_sums`: rel 1e-5 (the two kernels add the reward terms in a different association).
`hl_fused_last_impl()` proves which form actually ran.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [("flat", 4096, True), ("flat", 4096, False), ("stairs", 4096 + 37, True), ("stairs", 16384, True),
         ("flat", 65536, False), ("recover", 2048, True), ("flat", 36, True), ("flat", 1, True)]


def _run(monkeypatch, impl, cfg, state, hf, noise, single, steps=2):
    from gpu_helpers import make_env
    from isaacgymloco_b200 import _lib as L
    monkeypatch.setenv("HL_FUSED_IMPL", impl)
    env = make_env(cfg, state, hf, None, noise)
    env.single_launch = single
    env.refresh_buffers()
    used = []
    for _ in range(steps):
        env.fused_pre_reset()
        used.append(int(L.lib.hl_fused_last_impl()))
        env.fused_post_reset(with_reset_zero=True)
        env.common_step_counter += 1
    torch.cuda.synchronize()
    return env, used


@pytest.mark.parametrize("task,n,single", CASES)
def test_persistent_kernel_equals_tiled(monkeypatch, task, n, single):
    from gpu_helpers import assert_close, assert_equal
    from isaacgymloco_b200 import config as C, synthetic as S
    cfg = C.aliengo(task, num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=7)
    noise = S.make_noise(n, seed=8)
    a, used_a = _run(monkeypatch, "persist", cfg, state, hf, noise, single)
    b, used_b = _run(monkeypatch, "tiled", cfg, state, hf, noise, single)
    assert used_b == [0, 0]
    if task in ("flat", "stairs"):
        assert used_a == [1, 1], "the persistent kernel did not take this shard"
    ka, kb = int(a._n_reset.item()), int(b._n_reset.item())
    assert ka == kb
    assert_equal(a._reset_ids[:ka], b._reset_ids[:kb], "env_ids")
    assert_equal(a._term_priv[:ka], b._term_priv[:kb], "termination_privileged_obs")
    assert_equal(a._term_amp[:ka], b._term_amp[:kb], "terminal_amp_states")
    sa, sb = a.snapshot(), b.snapshot()
    for k in sa:
        if k in ("rew_buf", "episode_sums"):
            assert_close(sa[k], sb[k], k)
        else:
            assert_equal(sa[k], sb[k], k)


def test_persistent_kernel_philox_equals_tiled(monkeypatch):
    """In-kernel Philox noise: the streams are keyed by (seed, step, global env id), not by the tile or the kernel form."""
    from gpu_helpers import assert_equal
    from isaacgymloco_b200 import config as C, synthetic as S
    n = 8192
    cfg = C.aliengo("flat", num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=11)
    a, used = _run(monkeypatch, "persist", cfg, state, hf, None, True)
    b, _ = _run(monkeypatch, "tiled", cfg, state, hf, None, True)
    assert used == [1, 1]
    assert_equal(a.obs_buf, b.obs_buf, "obs_buf")
    assert_equal(a.privileged_obs_buf, b.privileged_obs_buf, "privileged_obs_buf")
    assert_equal(a._term_priv[:int(a._n_reset.item())], b._term_priv[:int(b._n_reset.item())], "termination_privileged_obs")


@pytest.mark.parametrize("task", ["flat", "stairs"])
def test_persistent_kernel_vs_oracle(monkeypatch, task):
    """The persistent form through the public post_physics_step against the oracle (pre-drawn noise, torch reset hooks)."""
    from gpu_helpers import assert_close, assert_equal, compare_snapshots, make_env
    from isaacgymloco_b200 import _lib as L, config as C, synthetic as S
    from oracle import torch_oracle as O
    n = 16384 + 24
    cfg = C.aliengo(task, num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=21)
    noise = S.make_noise(n, seed=22)
    targets = S.make_reset_targets(cfg, state, hf, seed=23)
    oenv = O.OracleEnv(cfg, S.to_device(state, "cuda"), hf.to("cuda"))
    oids, oterm, oamp = oenv.post_physics_step(S.to_device(noise, "cuda"), S.to_device(targets, "cuda"))
    monkeypatch.setenv("HL_FUSED_IMPL", "persist")
    env = make_env(cfg, state, hf, targets, noise)
    ids, term_obs, term_amp = env.post_physics_step()
    assert int(L.lib.hl_fused_last_impl()) == 1
    assert len(oids) > 100
    assert_equal(ids, oids, "env_ids")
    assert_close(term_obs, oterm, "termination_privileged_obs")
    assert_close(term_amp, oamp, "terminal_amp_states")
    compare_snapshots(env.snapshot(), oenv.snapshot())
