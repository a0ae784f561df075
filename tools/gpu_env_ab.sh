#!/bin/bash
# A/B of an env knob (dev): HL_AB_VAR=name HL_AB_VALS="a b c"
mkdir -p gpurun_out
for v in $HL_AB_VALS; do
  env $HL_AB_VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-latency > gpurun_out/ab.log 2>&1
  echo "$HL_AB_VAR=$v $(tail -1 gpurun_out/ab.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused_ms', round(d['roofline']['avg_launch_ms'],4), 'graph', round(d['cuda_graph']['ms_per_step'],3))" 2>&1 | tail -1)"
done
