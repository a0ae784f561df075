"""Dev tool: CUDA-event time of each kernel of one env-step in steady state (65,536 envs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

n = int(os.environ.get("N", 65536))
wl = bench.Workload(n, 24, 0, 1, torch.device("cuda", 0))
env = wl.env
for _ in range(30):
    wl.env_step()
torch.cuda.synchronize()
names = ["torque x4", "fused", "select+terminal", "fix-up"]
acc = [0.0] * 4
iters = 100
evs = []
for _ in range(iters):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record()
    for k in range(env.cfg_hot.decimation):
        env._compute_torques_into(env.delayed_actions[:, k], env.torques)
    e[1].record()
    import ctypes
    from isaacgymloco_b200 import _lib as L
    bufs = env._buffers()
    c, b = ctypes.byref(env._c), ctypes.byref(bufs)
    L.check(L.lib.hl_post_physics_fused(c, b, env.num_envs, L.stream()))
    e[2].record()
    if not env.single_launch:
        L.check(L.lib.hl_select_and_terminal(c, b, L.ptr(env._noise.get("term45")), L.ptr(env._noise.get("term187")),
                                             L.ptr(env._reset_ids), L.ptr(env._n_reset), L.ptr(env._term_priv),
                                             L.ptr(env._term_amp), L.ptr(env._selterm_ws), env.num_envs, L.stream()))
    e[3].record()
    env.fused_post_reset(with_reset_zero=True)
    e[4].record()
    env.common_step_counter += 1
    evs.append(e)
torch.cuda.synchronize()
for e in evs:
    for i in range(4):
        acc[i] += e[i].elapsed_time(e[i + 1])
print("resets per step:", int(env._n_reset.item()))
for nm, a in zip(names, acc):
    print(f"{nm:18s} {1e3 * a / iters:8.2f} us")
