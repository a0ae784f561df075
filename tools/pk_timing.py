"""Experiment: per-tile timestamps of the persistent fused kernel (build with
HL_LIB_NAME=libhimloco_b200_timing.so HL_DEFINES=-DHL_PK_TIMING; run with the same HL_LIB_NAME and
HL_FUSED_IMPL=persist).  Prints per-phase averages over tiles, by position in the CTA's sequence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from isaacgymloco_b200 import config as C, synthetic as S
from gpu_helpers import make_env

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = C.aliengo("flat", num_envs=n)
hf = S.make_terrain(cfg, seed=1)
state = S.make_state(cfg, n, hf, seed=7)
e = make_env(cfg, state, hf)
e.single_launch = (len(sys.argv) > 2 and sys.argv[2] == "compact")
e.refresh_buffers()
for _ in range(4):
    e.fused_pre_reset()
    e.fused_post_reset(with_reset_zero=True)
    e.common_step_counter += 1
torch.cuda.synchronize()
e._base_heights.zero_()
e.fused_pre_reset()
torch.cuda.synchronize()
t = e._base_heights.cpu().numpy()
nt = n // 32
t = t[:nt * 16].reshape(nt, 16)
names = ["S top", "S free", "S ticket", "S in_full", "S A done", "S base rdy", "S B done", "S lookback", "H start", "H hook0", "H hook1",
         "H envs done", "H free"]
print("kernel end (max timestamp) us:", t[:, :13].max())
order = np.argsort(t[:, 2])
for lo, hi, tag in [(0, 296, "first round"), (296, 888, "second"), (888, 1480, "third+"), (1480, nt, "last")]:
    sel = order[lo:hi]
    if len(sel) == 0:
        continue
    m = t[sel].mean(axis=0)
    print(f"--- tiles by ticket time rank {lo}..{hi} ({tag}): mean timestamps")
    print("  " + "  ".join(f"{names[k]}={m[k]:.1f}" for k in range(13)))
    d = t[sel]
    print(f"  S: wait free {np.mean(d[:,1]-d[:,0]):.2f}  ticket {np.mean(d[:,2]-d[:,1]):.2f}  load {np.mean(d[:,3]-d[:,2]):.2f}  A {np.mean(d[:,4]-d[:,3]):.2f}  "
          f"wait base {np.mean(d[:,5]-d[:,4]):.2f}  B {np.mean(d[:,6]-d[:,5]):.2f}  lookback {np.mean(d[:,7]-d[:,6]):.2f}  total {np.mean(d[:,7]-d[:,0]):.2f}")
    print(f"  H: pre-hook {np.mean(d[:,9]-d[:,8]):.2f}  hook {np.mean(d[:,10]-d[:,9]):.2f}  rest {np.mean(d[:,11]-d[:,10]):.2f}  tail {np.mean(d[:,12]-d[:,11]):.2f}  "
          f"total {np.mean(d[:,12]-d[:,8]):.2f}   S-B-done minus H-start {np.mean(d[:,6]-d[:,8]):.2f}")
