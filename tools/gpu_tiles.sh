#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
for t in 64 52 32 0; do
  HL_FUSED_EPB=$t python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_t$t.log 2>&1
  echo "tile=$t $(tail -1 gpurun_out/bench_t$t.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'graph ms', round(d['cuda_graph']['ms_per_step'],3), 'value', round(d['value']/1e6,1), 'lat4096 us', round(d['latency_4096']['us_per_env_step_graph'],1))")"
done
