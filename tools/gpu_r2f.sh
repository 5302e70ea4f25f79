#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv.py -m gpu -q -k "patch_kernel" > gpurun_out/pytest_patch.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_patch.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_patch.log | tail -8
for sk in 15 0; do for tr in 1 2; do echo "== JPB_CONV_SKIP=$sk TR=$tr"; for L in "layout layer1" "layout layer2"; do JPB_CONV_SKIP=$sk JPB_CONV_PATCH_TR=$tr timeout 120 python tools/bench_conv.py "$L" 10 2>&1 | cut -c1-240; done; done; done
