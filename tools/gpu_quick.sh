#!/bin/bash
# dev: GPU parity tests + a short bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu ${BENCH_FLAGS:---no-e2e} > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']/1e6,1), 'fused_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'graph', d.get('cuda_graph',{}).get('ms_per_step'), 'lat', d.get('latency_4096',{}).get('us_per_env_step_graph'))" 2>&1 | tail -1
