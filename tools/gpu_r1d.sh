#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])
for k,v in d['kernels'].items(): print(k, v)
"
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -45 gpurun_out/profile_step.txt
