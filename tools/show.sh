#!/bin/bash
TAG=${1:-rX}
python tools/ncu_summary.py gpurun_out/fused_$TAG.ncu-rep 2>&1 | head -${2:-45}
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','ms_per_step','cuda_graph','latency_4096'):
    print(k, d.get(k))
print('e2e',d['e2e']['value']); print('fused ms',d['roofline']['avg_launch_ms'], 'frac',d['roofline']['frac'], 'whole',d['roofline']['whole_step_frac'])"
