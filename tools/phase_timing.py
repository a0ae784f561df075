"""Dev tool: per-phase clock64 marks of the fused kernel (build with HL_DEFINES=-DHL_EXP_TIMING)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgymloco_b200 import config, synthetic
from isaacgymloco_b200.legged_robot import FusedLeggedRobot

n = int(os.environ.get("N", 65536))
epb = int(os.environ.get("HL_FUSED_EPB", 64))
cfg = config.aliengo("flat", num_envs=n)
hf = synthetic.make_terrain(cfg, seed=1)
st = synthetic.make_state(cfg, n, hf, seed=5)
env = FusedLeggedRobot(cfg, st, hf, device="cuda:0", seed=1)
for _ in range(5):
    env.fused_pre_reset()
torch.cuda.synchronize()
mh = env.measured_heights.view(n, -1)
rows = mh[::epb, :16].double()
names = ["p0 done", "A done(arrive2)", "term+stores", "after sync1", "reward done", "final sync", "phase2 done", "-",
         "scan setup", "base done", "first gathers", "after sync2", "scan loop done"]
mean = rows.mean(0); mx = rows.max(0).values; mn = rows.min(0).values
clk = 1.92e3  # cycles per us (approx)
for i, nm in enumerate(names):
    if nm == "-": continue
    print(f"{i:2d} {nm:18s} mean {mean[i]/clk:7.2f} us  min {mn[i]/clk:7.2f}  max {mx[i]/clk:7.2f}")
