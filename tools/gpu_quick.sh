#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv.py tests/test_model_parity.py tests/test_heads.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED" | tail -4
for se in 0 1; do
JPB_SPLIT_EPILOGUE=$se timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_se$se.json 2> gpurun_out/bench_se$se.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_se$se.json').read())
print('split_epilogue $se', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if 'conv' in k})
"
done
