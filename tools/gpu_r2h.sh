#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_h.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2), "conv frac", round(d["roofline"]["frac"],4), "photo frac", round(d["roofline_photometric"]["frac"],4))
    print(" gpu_eager", d.get("gpu_eager_baseline")); print(" cpu", d.get("cpu_baseline")); print(" launches", d["gpu_launches"]/d["steps"], d["kernel_timing"])
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_h.err").read()[-2500:])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
for c in C3 C4 C5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", d["config"]["workload"][:90], round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms", "gpu_eager", (d.get("gpu_eager_baseline") or {}).get("value"))
except Exception as e:
    print("bench $c unreadable", e); print(open("gpurun_out/bench_$c.err").read()[-1500:])
PY
done
bash tools/gpu_traffic.sh
