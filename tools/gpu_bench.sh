#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; tail -3 gpurun_out/bench_eager.err
python -c "
import json
for f in ('graph','eager'):
    try:
        d=json.loads(open('gpurun_out/bench_%s.json'%f).read())
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
d=json.loads(open('gpurun_out/bench_graph.json').read())
for k,v in d['kernels'].items(): print(k, v)
"
