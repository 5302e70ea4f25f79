#!/bin/bash
# round-end check, 1 GPU: full gpu suite, smoke, bench line (with cpu_baseline), conv + photometric micro-benchmarks
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_full.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline_photometric']['frac'], d['cpu_baseline']['value'])
"
timeout 200 python tools/bench_conv.py > gpurun_out/conv_microbench_final.jsonl 2>/dev/null
timeout 100 python tools/bench_photometric.py --B 4 > gpurun_out/photo_b4.json 2>/dev/null; cut -c1-260 gpurun_out/photo_b4.json
timeout 200 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -3 gpurun_out/profile_step.txt | tail -2
