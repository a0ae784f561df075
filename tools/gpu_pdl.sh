#!/bin/bash
# A/B of programmatic dependent launch (dev)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env.py -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
for p in 1 0 1; do
  HL_PDL=$p python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_pdl$p.log 2>&1
  echo "PDL=$p $(tail -1 gpurun_out/bench_pdl$p.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']/1e6,1), 'fused_ms', round(d['roofline']['avg_launch_ms'],4), 'direct ms', d.get('direct_ms_per_step'), 'graph', d.get('cuda_graph'), 'lat', d.get('latency_4096'), 'e2e', round(d['e2e']['value']/1e6,2))" 2>&1 | tail -1)"
done
