#!/bin/bash
# A/B the fused kernel occupancy knob (tests once, then two quick benches)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
for occ in 3 4; do
  HL_FUSED_OCC=$occ python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-latency > gpurun_out/bench_occ$occ.log 2>&1
  echo "occ=$occ $(tail -1 gpurun_out/bench_occ$occ.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d.get('cuda_graph'))")"
done
