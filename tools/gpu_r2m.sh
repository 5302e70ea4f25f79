#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv.py tests/test_zzz_conv_multitile.py tests/test_model_parity.py -m gpu -q > gpurun_out/pytest_m.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_m.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_m.log | tail -12
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_m.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4), "conv ms", round(d["roofline"]["ms_per_step"],2))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_m.err").read()[-1500:])
PY
timeout 200 python tools/bench_conv.py "stem" 10 2>&1 | cut -c1-300
