#!/bin/bash
mkdir -p gpurun_out
export JPB_CONV_VARIANT=2
timeout 300 python -m pytest tests/test_conv.py -m gpu -q -x 2>&1 | tail -3
for cfg in "cb 0"; do
  set -- $cfg
  echo "=== korder $1 l1 $2"
  JPB_CONV_KORDER=$1 JPB_CONV_L1=$2 timeout 300 python tools/bench_conv.py > gpurun_out/convvar_$1_$2.jsonl 2> gpurun_out/convvar_$1_$2.err || tail -5 gpurun_out/convvar_$1_$2.err
  python - <<PY
import json
for l in open('gpurun_out/convvar_$1_$2.jsonl'):
    r=json.loads(l); print("%-42s fwd %7.1f us %6.1f TF | dgrad %7.1f us | wgrad %7.1f us %6.1f TF" % (r['layer'][:42], r['fwd_ms']*1e3, r['fwd_tflops'], r.get('dgrad_ms',0)*1e3, r['wgrad_ms']*1e3, r['wgrad_tflops']))
PY
done
