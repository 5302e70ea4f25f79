#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_losses.py tests/test_zz_photometric_kept.py -m gpu -q -k "photometric" > gpurun_out/pytest_y.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_y.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_y.log | tail -8 | cut -c1-300
for dbg in "" "--debug-outputs"; do timeout 120 python tools/bench_photometric.py --B 4 --variant 3 $dbg 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('dbg', d['debug_outputs'], 'fwd ms %.4f (frac %.3f)  bwd ms %.4f (frac %.3f)' % (d['fwd_ms'], d['fwd_frac'], d['bwd_ms'], d['bwd_frac']))"; done
timeout 300 python -m pytest tests/test_model_parity.py -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_y.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4), "photo", round(d["roofline_photometric"]["ms_per_launch"]*1e3,1), "us frac", round(d["roofline_photometric"]["frac"],3), "bwd", round(d["roofline_photometric"]["backward"]["ms_per_launch"]*1e3,1), "us frac", round(d["roofline_photometric"]["backward"]["frac"],3))
    for k, v in d["kernels"].items():
        if "pool" in k: print("  %-18s %8.3f ms/step  n=%4d  %8.2f us" % (k, v["ms_per_step"], v["launches_per_step"], v["ms_per_launch"]*1e3))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_y.err").read()[-1500:])
PY
