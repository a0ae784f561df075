#!/bin/bash
# round-2 closing run on one B200: tests, smoke, both bench arms, launch list, one full ncu capture, sanitizers
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > gpurun_out/r2_final_gpu_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>> gpurun_out/r2_final_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 120 --csv --log-file gpurun_out/r2_final_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu --no-latency > gpurun_out/r2_final_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:hl_post_physics_fused -s 30 -c 1 -f -o gpurun_out/fused_r2final \
  python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu --no-latency > gpurun_out/r2_final_ncu.log 2>&1
for tool in memcheck synccheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_reset.py -m gpu -q -x --timeout 140 \
    > gpurun_out/r2_final_$tool.log 2>&1; echo "exit $?" >> gpurun_out/r2_final_$tool.log
done
HL_FUSED_IMPL=persist timeout 120 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/persist_smoke.py > gpurun_out/r2_final_synccheck_persist.log 2>&1; echo "exit $?" >> gpurun_out/r2_final_synccheck_persist.log
tail -3 gpurun_out/r2_final_gpu_tests.log gpurun_out/r2_final_smoke.log gpurun_out/r2_final_memcheck.log gpurun_out/r2_final_synccheck.log gpurun_out/r2_final_synccheck_persist.log
