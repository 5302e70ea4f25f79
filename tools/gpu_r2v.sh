#!/bin/bash
# round-2 evidence: whole GPU suite, smoke, bench line (all legs), launch list of the bench command, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_v.log
grep -E "passed|failed|FAILED|ERROR|pytest exit" gpurun_out/pytest_gpu_v.log | tail -15 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v.log 2>&1; tail -2 gpurun_out/smoke_v.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_v.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"],2), " conv frac", round(d["roofline"]["frac"],4), "photo frac", round(d["roofline_photometric"]["frac"],4), "cpu", d.get("cpu_baseline"), "eager", d.get("gpu_eager_baseline"))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_v.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-graph > gpurun_out/ncu_bench_v.log 2>&1
wc -l gpurun_out/launches_v.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd_kernel -c 1 -o gpurun_out/ncu_conv_rows_fwd -f \
  python tools/bench_conv.py merge1 1 > gpurun_out/ncu_conv_rows_fwd.log 2>&1
ncu -i gpurun_out/ncu_conv_rows_fwd.ncu-rep --page raw --csv > gpurun_out/ncu_conv_rows_fwd_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad_kernel -c 1 -o gpurun_out/ncu_conv_rows_wgrad -f \
  python tools/bench_conv.py merge1 1 > gpurun_out/ncu_conv_rows_wgrad.log 2>&1
ncu -i gpurun_out/ncu_conv_rows_wgrad.ncu-rep --page raw --csv > gpurun_out/ncu_conv_rows_wgrad_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
