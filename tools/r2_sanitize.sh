#!/bin/bash
# compute-sanitizer over the -m gpu suite (memcheck, synccheck: all tests; racecheck: the round-2 kernels' files)
for tool in memcheck synccheck; do
  timeout 170 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests -m gpu -q --timeout 160 > gpurun_out/r2_san_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r2_san_$tool.log
done
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_reset.py tests/test_gpu_persist.py -m gpu -q --timeout 190 > gpurun_out/r2_san_racecheck.log 2>&1
echo "exit $?" >> gpurun_out/r2_san_racecheck.log
tail -n 4 gpurun_out/r2_san_*.log
