#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv.py tests/test_losses.py -m gpu -q -x > gpurun_out/conv_test.log 2>&1; tail -3 gpurun_out/conv_test.log


python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read())
print(d['value'], d['ms_per_step'])
for k,v in d['kernels'].items(): print(k, v)
"
