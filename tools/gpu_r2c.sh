#!/bin/bash
mkdir -p gpurun_out
for c in 1.00067702 1.0 1.001; do JPB_TF32_TRUNC_COMP=$c timeout 300 python -m pytest tests/test_model_parity.py -m gpu -q -s -k calibrated > gpurun_out/calib_$c.log 2>&1; echo "== TRUNC_COMP=$c"; grep -E "passed|failed" gpurun_out/calib_$c.log | head -5; grep -A 22 "TF32 calibration" gpurun_out/calib_$c.log | head -24; done
