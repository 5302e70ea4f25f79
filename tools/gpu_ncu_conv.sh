#!/bin/bash
mkdir -p gpurun_out
export JPB_CONV_VARIANT=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 3 -c 1 -o gpurun_out/ncu_conv_v2_merge -f python tools/bench_conv.py "merge1" 2 > gpurun_out/ncu_conv_v2.log 2>&1
ls -la gpurun_out/*.ncu-rep
