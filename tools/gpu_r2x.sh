#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_losses.py -m gpu -q -k "maxpool" > gpurun_out/pytest_x.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_x.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_x.log | tail -6 | cut -c1-300
timeout 200 python tools/bench_misc.py 2>&1 | grep pool | cut -c1-250
timeout 300 python -m pytest tests/test_model_parity.py -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4))
    for k, v in d["kernels"].items():
        if "pool" in k: print("  %-18s %8.3f ms/step  n=%4d  %8.2f us" % (k, v["ms_per_step"], v["launches_per_step"], v["ms_per_launch"]*1e3))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_x.err").read()[-1500:])
PY
