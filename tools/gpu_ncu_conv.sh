#!/bin/bash
mkdir -p gpurun_out
export JPB_CONV_VARIANT=2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 3 -c 1 -o gpurun_out/ncu_conv_l1 -f python tools/bench_conv.py "layout layer1" 2 > gpurun_out/ncu_conv_l1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 3 -c 1 -o gpurun_out/ncu_conv_ic1 -f python tools/bench_conv.py "iconv1" 2 > gpurun_out/ncu_conv_ic1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad -s 3 -c 1 -o gpurun_out/ncu_conv_wg -f python tools/bench_conv.py "iconv1" 2 > gpurun_out/ncu_conv_wg.log 2>&1
ls -la gpurun_out/*.ncu-rep
