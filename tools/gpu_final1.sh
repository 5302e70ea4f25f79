#!/bin/bash
# round-end evidence, 1 GPU: full bench line (with cpu_baseline), reference arm, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; tail -2 gpurun_out/bench_n1_full.err; cut -c1-400 gpurun_out/bench_n1_full.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_step.csv
