#!/bin/bash
# CTA-pair (cta_group::2) forward kernel: parity (multi-tile + layer suite), then A/B of the wide layers, then the bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zzz_conv_multitile.py -m gpu -q -x > gpurun_out/pytest_r.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_r.log | tail -8 | cut -c1-300
if grep -q "pytest exit 0" gpurun_out/pytest_r.log; then
  for pair in 0 1 2; do
    for l in iconv1 merge1 crp1 iconv2; do
      JPB_CONV_PAIR=$pair timeout 120 python tools/bench_conv.py "$l" 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('pair=$pair %-42s fwd %7.1f us %5.0f TF | dgrad %7.1f us %5.0f | wgrad %7.1f us %5.0f'%(r['layer'][:42], r['fwd_ms']*1e3, r['fwd_tflops'], r.get('dgrad_ms',0)*1e3, r.get('dgrad_tflops',0), r['wgrad_ms']*1e3, r['wgrad_tflops']))"
    done
  done
  timeout 600 python -m pytest tests/test_conv.py tests/test_model_parity.py -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2.log
  grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_r2.log | tail -8 | cut -c1-300
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4), "conv ms", round(d["roofline"]["ms_per_step"],2))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_r.err").read()[-1500:])
PY
fi
