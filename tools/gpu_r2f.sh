#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1 2 4 7 15; do echo "== JPB_CONV_SKIP=$sk"; for L in "layout layer1" "layout layer2" "layout layer3"; do JPB_CONV_SKIP=$sk timeout 120 python tools/bench_conv.py "$L" 10 2>&1 | cut -c1-240; done; done
echo "== gather kernels, skip 0 / 1 (no A) / 2 (no B) / 3"; for sk in 0 1 2 3; do JPB_CONV_SKIP=$sk timeout 120 python tools/bench_conv.py "merge1" 10 2>&1 | cut -c1-200; done
