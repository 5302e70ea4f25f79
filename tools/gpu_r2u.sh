#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv.py tests/test_zzz_conv_multitile.py -m gpu -q -x > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_u.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_u.log | tail -12 | cut -c1-300
if grep -q "pytest exit 0" gpurun_out/pytest_u.log; then
for rows in 0 1; do
  for l in iconv1 merge1 crp1 iconv2; do
    JPB_FWD_ROWS=$rows timeout 120 python tools/bench_conv.py "$l" 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('rows=$rows %-42s fwd %7.1f us %5.0f TF | dgrad %7.1f us %5.0f | wgrad %7.1f us %5.0f'%(r['layer'][:42], r['fwd_ms']*1e3, r['fwd_tflops'], r.get('dgrad_ms',0)*1e3, r.get('dgrad_tflops',0), r['wgrad_ms']*1e3, r['wgrad_tflops']))"
  done
done
timeout 600 python -m pytest tests/test_model_parity.py -m gpu -q > gpurun_out/pytest_u2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_u2.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_u2.log | tail -8 | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_u.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4), "conv ms", round(d["roofline"]["ms_per_step"],2))
    for k, v in d["kernels"].items(): print("  %-18s %8.3f ms/step  n=%4d  %8.2f us" % (k, v["ms_per_step"], v["launches_per_step"], v["ms_per_launch"]*1e3))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_u.err").read()[-1500:])
PY
fi
