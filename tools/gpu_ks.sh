#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_conv.py -m gpu -x -q 2>&1 | tail -2
for sl in 148 296; do
JPB_KSPLIT_SLOTS=$sl timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ks$sl.json 2> gpurun_out/bench_ks$sl.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_ks$sl.json').read())
print('slots $sl', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if 'conv' in k})
"
done
