#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_trainer.py -m gpu -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_d.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_gpu_d.log | tail -20
for ws in 1; do
  JPB_WGRAD_STREAMS=$ws timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ws$ws.json 2> gpurun_out/bench_ws$ws.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ws$ws.json").read().strip().splitlines()[-1])
    print("wgrad_streams=$ws", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2))
except Exception as e:
    print("bench ws=$ws unreadable", e); print(open("gpurun_out/bench_ws$ws.err").read()[-1500:])
PY
done
