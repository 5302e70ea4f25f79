#!/bin/bash
# round 2, session 2, call 1: per-kernel device times with back-pressure, CUPTI table of one eager step, launch list
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4), "conv ms", round(d["roofline"]["ms_per_step"],2), "bound frac", round(d["roofline"]["frac_of_per_kernel_bound"],3))
    for k, v in d["kernels"].items(): print("  %-18s %8.3f ms/step  n=%4d  %8.2f us" % (k, v["ms_per_step"], v["launches_per_step"], v["ms_per_launch"]*1e3))
    print(d["kernel_timing"])
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_n.err").read()[-1500:])
PY
timeout 300 python tools/profile_step.py > gpurun_out/step_kernel_times_cupti.txt 2>&1; head -50 gpurun_out/step_kernel_times_cupti.txt
