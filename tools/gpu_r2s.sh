#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv.py -m gpu -q -x -k "tma_rows" > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_s.log | tail -12 | cut -c1-300
for rows in 0 1 2; do
  for l in iconv1 merge1 crp1 "layer1 64" "layout layer1" "layer3 256" "layout layer3"; do
    JPB_WGRAD_ROWS=$rows timeout 120 python tools/bench_conv.py "$l" 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('rows=$rows %-42s wgrad %7.1f us %5.0f TF'%(r['layer'][:42], r['wgrad_ms']*1e3, r['wgrad_tflops']))"
  done
done
