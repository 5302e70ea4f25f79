#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 4; do
export JPB_CONV_VARIANT=$v
echo "=== variant $v"
timeout 120 python -m pytest tests/test_conv.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python tools/bench_conv.py > gpurun_out/convvar_$v.jsonl 2> gpurun_out/convvar_$v.err || tail -5 gpurun_out/convvar_$v.err
  python - <<PY
import json
for l in open('gpurun_out/convvar_$v.jsonl'):
    r=json.loads(l); print("%-42s fwd %7.1f us %6.1f TF | dgrad %7.1f us | wgrad %7.1f us %6.1f TF" % (r['layer'][:42], r['fwd_ms']*1e3, r['fwd_tflops'], r.get('dgrad_ms',0)*1e3, r['wgrad_ms']*1e3, r['wgrad_tflops']))
PY
done
