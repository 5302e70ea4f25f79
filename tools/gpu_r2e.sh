#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv.py tests/test_model_parity.py tests/test_zzz_conv_multitile.py -m gpu -q > gpurun_out/pytest_patch.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_patch.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_patch.log | tail -20
for pt in 0 1; do
  JPB_CONV_PATCH=$pt timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_patch$pt.json 2> gpurun_out/bench_patch$pt.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_patch$pt.json").read().strip().splitlines()[-1])
    print("patch=$pt", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2), "conv frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("bench patch=$pt unreadable", e); print(open("gpurun_out/bench_patch$pt.err").read()[-1500:])
PY
done
JPB_BRANCH_STREAMS=0 timeout 300 python tools/conv_layers.py > gpurun_out/conv_layers_patch.txt 2>&1; head -40 gpurun_out/conv_layers_patch.txt
