#!/bin/bash
# A/B experiments (dev): "defines;ENV=val ..." sets separated by |
mkdir -p gpurun_out
IFS='|' read -ra SETS <<< "${HL_EXP_SETS:-;}"
for set in "${SETS[@]}"; do
  d="${set%%;*}"; envs="${set#*;}"
  HL_DEFINES="$d" python isaacgymloco_b200/build.py --force > gpurun_out/build.log 2>&1 || { echo "[$set] build failed"; tail -3 gpurun_out/build.log; continue; }
  if [ -n "$HL_EXP_TESTS" ]; then env $envs python -m pytest tests/test_gpu_env.py -m gpu -q -x > gpurun_out/pytest_exp.log 2>&1; t=$(tail -1 gpurun_out/pytest_exp.log); else t=notest; fi
  env $envs python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-latency > gpurun_out/exp.log 2>&1
  echo "[$set] $t | $(tail -1 gpurun_out/exp.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'ms_per_step', round(d['ms_per_step'],3), 'graph', round(d['cuda_graph']['ms_per_step'],3) if 'ms_per_step' in d.get('cuda_graph',{}) else d.get('cuda_graph'))" 2>&1 | tail -1)"
done
python isaacgymloco_b200/build.py --force > /dev/null 2>&1
