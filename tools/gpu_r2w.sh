#!/bin/bash
# other BASELINE configurations on one GPU with the round-2 kernels (no baselines: they are in r2_bench_n1_final.json / r2_scaling_n8.md)
mkdir -p gpurun_out
for c in C3 C4 C5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_w_$c.json 2> gpurun_out/bench_w_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_w_$c.json").read().strip().splitlines()[-1])
    print("$c", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"],2), "conv frac", round(d["roofline"]["frac"],4), d["config"]["workload"][:90])
except Exception as e:
    print("$c bench unreadable", e); print(open("gpurun_out/bench_w_$c.err").read()[-1200:])
PY
done
