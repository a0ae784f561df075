#!/bin/bash
# dev loop on the GPU box: tests, bench, launch list, one full ncu capture of the fused kernel
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu --no-latency > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hl_post_physics_fused -s 30 -c 1 -o gpurun_out/fused_$TAG python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu --no-latency > gpurun_out/b_ncu2.log 2>&1
