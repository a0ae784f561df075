#!/bin/bash
# A/B experiments on the fused kernel (dev): rebuild with -D knobs, run tests once, time the bench
mkdir -p gpurun_out
IFS='|' read -ra SETS <<< "${HL_EXP_SETS:-}"
for d in "${SETS[@]}"; do
  HL_DEFINES="$d" python isaacgymloco_b200/build.py --force > gpurun_out/build.log 2>&1 || { echo "[$d] build failed"; tail -3 gpurun_out/build.log; continue; }
  regs=$(grep -A2 "fused_kernelILb0ELi6" gpurun_out/build.log | grep -o "Used [0-9]* registers" | head -1)
  t=skip
  python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-latency > gpurun_out/exp.log 2>&1
  echo "[$d] $regs | $t | $(tail -1 gpurun_out/exp.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'ms_per_step', round(d['ms_per_step'],3), 'graph', round(d['cuda_graph']['ms_per_step'],3))")"
done
python isaacgymloco_b200/build.py --force > /dev/null 2>&1
