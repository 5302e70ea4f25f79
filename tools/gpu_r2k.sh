#!/bin/bash
# photometric forward (variant 3): 2 vs 3 resident blocks per SM (80 registers, no spills at 3)
mkdir -p gpurun_out
for b in 4; do timeout 120 python tools/bench_photometric.py --B $b --variant 3 > gpurun_out/photo_mb2_b$b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/photo_mb2_b$b.json')); print('minblocks 2 B', d['B'], 'fwd ms', round(d['fwd_ms'],4), 'bwd ms', round(d['bwd_ms'],4))"; done
rm -rf /tmp/jpb_exp && mkdir -p /tmp/jpb_exp && cp -r bench.py tools tests include jperceiver_b200 oracle mono mmcv MEASURED_PEAKS.json /tmp/jpb_exp/ 2>/dev/null
( cd /tmp/jpb_exp && sed -i 's/__global__ void __launch_bounds__(256, 2) photometric_fwd_kernel(JpbPhotoArgs a) {/__global__ void __launch_bounds__(256, 3) photometric_fwd_kernel(JpbPhotoArgs a) {/' jperceiver_b200/csrc/photometric.cu && rm -f jperceiver_b200/csrc/libjpb200.so && python -m jperceiver_b200.build > /dev/null 2>&1 \
  && for b in 4; do timeout 120 python tools/bench_photometric.py --B $b --variant 3 > $OLDPWD/gpurun_out/photo_mb3_b$b.json 2>/dev/null; python -c "
import json; d=json.load(open('$OLDPWD/gpurun_out/photo_mb3_b$b.json')); print('minblocks 3 B', d['B'], 'fwd ms', round(d['fwd_ms'],4), 'bwd ms', round(d['bwd_ms'],4))"; done )
