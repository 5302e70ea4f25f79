#!/bin/bash
# per-layer conv table with back-pressure timing; ncu --set full of the photometric backward and forward (scale 0 and scale 3)
mkdir -p gpurun_out
timeout 300 python tools/conv_layers.py > gpurun_out/conv_layers.txt 2>&1; head -40 gpurun_out/conv_layers.txt | cut -c1-220
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric_bwd -s 8 -c 4 -o gpurun_out/ncu_photo_bwd -f \
  python tools/bench_photometric.py --B 4 --iters 2 > gpurun_out/ncu_photo_bwd.log 2>&1
ncu -i gpurun_out/ncu_photo_bwd.ncu-rep --page raw --csv > gpurun_out/ncu_photo_bwd_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric_fwd -s 8 -c 4 -o gpurun_out/ncu_photo_fwd -f \
  python tools/bench_photometric.py --B 4 --iters 2 > gpurun_out/ncu_photo_fwd.log 2>&1
ncu -i gpurun_out/ncu_photo_fwd.ncu-rep --page raw --csv > gpurun_out/ncu_photo_fwd_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
