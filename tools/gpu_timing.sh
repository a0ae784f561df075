#!/bin/bash
mkdir -p gpurun_out
HL_DEFINES="-DHL_EXP_TIMING" python isaacgymloco_b200/build.py --force > gpurun_out/build.log 2>&1 || { tail -5 gpurun_out/build.log; exit 1; }
for e in 64 52; do echo "== tile $e"; HL_FUSED_EPB=$e timeout 300 python tools/phase_timing.py; done
python isaacgymloco_b200/build.py --force > /dev/null 2>&1
