#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv.py -m gpu -q > gpurun_out/conv_test.log 2>&1; echo "exit $?" >> gpurun_out/conv_test.log
tail -30 gpurun_out/conv_test.log
