#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_gpu_g.log | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2), "conv frac", round(d["roofline"]["frac"],4))
    for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:14]: print('  %-22s %8.3f ms/step n=%4d  %8.2f us each'%(k,v['ms_per_step'],v['launches_per_step'],v['ms_per_launch']*1e3))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_g.err").read()[-1500:])
PY
JPB_BRANCH_STREAMS=0 JPB_WGRAD_STREAMS=0 timeout 300 python tools/conv_layers.py > gpurun_out/conv_layers_g.txt 2>&1; head -32 gpurun_out/conv_layers_g.txt
