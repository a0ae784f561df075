"""GPU smoke for the persistent fused kernel: a few shapes / modes, each checked against the tiled
kernel (HL_FUSED_IMPL=tiled) bit for bit, progress printed as it goes (run under `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from isaacgymloco_b200 import config as C, synthetic as S
from gpu_helpers import make_env


def run(task, n, single, steps=2, noise=True):
    cfg = C.aliengo(task, num_envs=n)
    hf = S.make_terrain(cfg, seed=1)
    state = S.make_state(cfg, n, hf, seed=7)
    nz = S.make_noise(n, seed=8) if noise else None
    envs = []
    for impl in ("persist", "tiled"):
        os.environ["HL_FUSED_IMPL"] = impl
        e = make_env(cfg, state, hf, None, nz)
        e.single_launch = single
        e.refresh_buffers()
        for _ in range(steps):
            e.fused_pre_reset()
            e.fused_post_reset(with_reset_zero=True)
            e.common_step_counter += 1
        torch.cuda.synchronize()
        envs.append(e)
    a, b = envs
    ka, kb = int(a._n_reset.item()), int(b._n_reset.item())
    bad = []
    if ka != kb or not torch.equal(a._reset_ids[:ka], b._reset_ids[:kb]):
        bad.append("ids")
    elif not torch.equal(a._term_priv[:ka], b._term_priv[:kb]):
        bad.append("term_priv")
    sa, sb = a.snapshot(), b.snapshot()
    for k in sa:
        if not torch.equal(sa[k], sb[k]):
            d = (sa[k].float() - sb[k].float()).abs().max().item()
            bad.append(f"{k}(max {d:.3g})")
    print(f"  {task} n={n} single={single} noise={noise}: resets={ka} {'OK' if not bad else 'DIFF ' + ' '.join(bad)}", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    for args in [("flat", 4096, True), ("flat", 4096, False), ("stairs", 4096 + 37, True), ("stairs", 16384, True),
                 ("flat", 65536, False), ("flat", 65536, True), ("recover", 2048, True), ("flat", 36, True), ("flat", 1, True)]:
        print("case", args, flush=True)
        run(*args)
    run("flat", 8192, True, noise=False)
    print(f"done in {time.time() - t0:.1f}s", flush=True)
