#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_losses.py tests/test_zz_photometric_kept.py -m gpu -q > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_q.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_q.log | tail -12 | cut -c1-300
timeout 120 tools/microbench/mma_ceiling > gpurun_out/mma_ceiling.txt 2>&1; cat gpurun_out/mma_ceiling.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric_bwd -s 8 -c 4 -o gpurun_out/ncu_photo_bwd4 -f \
  python tools/bench_photometric.py --B 4 --iters 2 --debug-outputs > gpurun_out/ncu_photo_bwd4.log 2>&1
ncu -i gpurun_out/ncu_photo_bwd4.ncu-rep --page raw --csv > gpurun_out/ncu_photo_bwd4_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_photo_bwd4.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/ncu_photo_bwd4_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric_fwd -s 8 -c 4 -o gpurun_out/ncu_photo_fwd4 -f \
  python tools/bench_photometric.py --B 4 --iters 2 --debug-outputs > gpurun_out/ncu_photo_fwd4.log 2>&1
ncu -i gpurun_out/ncu_photo_fwd4.ncu-rep --page raw --csv > gpurun_out/ncu_photo_fwd4_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_photo_fwd4.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/ncu_photo_fwd4_source.csv 2>/dev/null
ls -la gpurun_out | tail -8
