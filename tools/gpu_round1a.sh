#!/bin/bash
# GPU session 1: loss-kernel parity on the real library, micro-benchmark, ncu evidence.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/bench_photometric.py --B 4 > gpurun_out/photo_b4.json 2> gpurun_out/photo_b4.err; cat gpurun_out/photo_b4.json
python tools/bench_photometric.py --B 8 > gpurun_out/photo_b8.json 2> gpurun_out/photo_b8.err; cat gpurun_out/photo_b8.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:photometric -s 8 -c 8 -o gpurun_out/photo_prof python tools/bench_photometric.py --B 4 --iters 2 > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/photo_launches.csv python tools/bench_photometric.py --B 4 --iters 2 > gpurun_out/ncu_launch.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_losses.py -m gpu -x -q -k "not full_size" > gpurun_out/memcheck.log 2>&1; tail -5 gpurun_out/memcheck.log
ls -la gpurun_out
