"""CPU, gloo, world_size 2: the host-side logic of the env-sharded path --
flat gradient all-reduce, advantage-moment all-reduce, AMP normaliser moment merge, global env
numbering of the shards.  (The GPU path uses the same code over NCCL.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from isaacgymloco_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    try:
        q.put((rank, globals()[fn_name](rank, world)))
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2, attempts=2):
    """Spawn `world` gloo ranks on a fresh port; a rendezvous that does not complete (the probed port can be taken
    between the probe and rank 0's bind) is torn down and retried once on another port."""
    import queue
    ctx = mp.get_context("spawn")
    for attempt in range(attempts):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            out = dict(q.get(timeout=90) for _ in range(world))
        except queue.Empty:
            for p in procs:
                p.terminate()
                p.join(timeout=10)
            if attempt + 1 == attempts:
                raise
            continue
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        return out


def _grad_case(rank, world):
    from isaacgymloco_b200 import dist as D
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
    red = D.FlatGradAllReducer(net.parameters())
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(64, 8, generator=g), torch.randn(64, 3, generator=g)
    lo, hi = D.shard_range(64, rank, world)
    red.zero_grad()
    torch.nn.functional.mse_loss(net(x[lo:hi]), y[lo:hi]).backward()
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in net.parameters())   # still views
    red.all_reduce_mean()
    norm = red.clip_grad_norm_(1e9)
    # single-process reference: the global-batch gradient
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
    ref.load_state_dict(net.state_dict())
    torch.nn.functional.mse_loss(ref(x), y).backward()
    want = torch.cat([p.grad.flatten() for p in ref.parameters()])
    kl = D.all_reduce_scalar_mean(torch.tensor(float(rank)))
    return red.flat.clone().numpy(), want.numpy(), float(norm), float(kl)


def test_flat_grad_allreduce_equals_global_batch_gradient():
    out = _run("_grad_case")
    for r in (0, 1):
        got, want, norm, kl = out[r]
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)
        assert abs(norm - np.linalg.norm(want)) < 1e-5 and kl == 0.5
    np.testing.assert_array_equal(out[0][0], out[1][0])


def _moments_case(rank, world):
    from isaacgymloco_b200.amp_discriminator import allreduce_moments
    from isaacgymloco_b200 import dist as D
    g = torch.Generator().manual_seed(9)
    data = torch.randn(1000, 30, generator=g, dtype=torch.float64) * 3 + 1
    adv = torch.randn(24, 512, 1, generator=g)
    lo, hi = D.shard_range(1000, rank, world)
    mine = data[lo:hi]
    gm, gv, cnt = allreduce_moments(mine.mean(0), mine.var(0, unbiased=False), float(hi - lo), "world")
    alo, ahi = D.shard_range(512, rank, world)
    a = adv[:, alo:ahi].double()
    m = torch.stack([a.sum(), (a * a).sum(), torch.tensor(float(a.numel()), dtype=torch.float64)])
    D.all_reduce_sum_(m)
    mean = m[0] / m[2]
    std = torch.sqrt((m[1] - m[0] * mean) / (m[2] - 1))
    return (gm.numpy(), gv.numpy(), cnt, data.mean(0).numpy(), data.var(0, unbiased=False).numpy(),
            float(mean), float(std), float(adv.double().mean()), float(adv.double().std()))


def test_sharded_statistics_equal_single_process():
    out = _run("_moments_case")
    for r in (0, 1):
        gm, gv, cnt, wm, wv, mean, std, wmean, wstd = out[r]
        assert cnt == 1000.0
        np.testing.assert_allclose(gm, wm, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(gv, wv, rtol=1e-10, atol=1e-12)
        assert abs(mean - wmean) < 1e-12 and abs(std - wstd) < 1e-12


def test_shard_range_uses_global_env_numbering():
    from isaacgymloco_b200 import dist as D
    from isaacgymloco_b200 import config as C
    assert D.shard_range(65536 * 8, 3, 8) == (3 * 65536, 4 * 65536)
    with pytest.raises(ValueError):
        D.shard_range(10, 0, 3)
    cfg = C.aliengo("stairs", num_envs=16384 * 2, env_id_offset=16384)
    c = cfg.to_c()
    assert c.env_id_offset == 16384 and c.stairsup_start == 6554 and c.stairsup_end == 16384
