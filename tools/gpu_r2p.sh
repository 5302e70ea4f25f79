#!/bin/bash
# photometric: new backward schedule (4) and identity terms shared across scales — parity suites, A/B timings, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_losses.py tests/test_zz_photometric_kept.py tests/test_zzy_photometric_packed.py tests/test_model_parity.py -m gpu -q > gpurun_out/pytest_p.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_p.log
grep -E "passed|failed|FAILED|ERROR|pytest exit|^E  " gpurun_out/pytest_p.log | tail -12
rm -f gpurun_out/photo_ab2.jsonl
for dbg in "" "--debug-outputs"; do
  for args in "--bwd-variant 1 --ident-every-scale" "--bwd-variant 4 --ident-every-scale" "--bwd-variant 4"; do
    timeout 120 python tools/bench_photometric.py --B 4 --variant 3 $args $dbg >> gpurun_out/photo_ab2.jsonl 2>> gpurun_out/photo_ab2.err
  done
done
timeout 120 python tools/bench_photometric.py --B 8 --variant 3 --debug-outputs >> gpurun_out/photo_ab2.jsonl 2>> gpurun_out/photo_ab2.err
python - <<PY
import json
for l in open("gpurun_out/photo_ab2.jsonl"):
    d = json.loads(l)
    print("B", d["B"], "dbg", d["debug_outputs"], "bwdv", d["bwd_variant"], d["identity_terms"], "fwd ms %.4f (frac %.3f)  bwd ms %.4f (frac %.3f)" % (d["fwd_ms"], d["fwd_frac"], d["bwd_ms"], d["bwd_frac"]))
PY
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_p.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  conv frac", round(d["roofline"]["frac"],4), "photo", d["roofline_photometric"]["ms_per_launch"], d["roofline_photometric"]["frac"], "bwd", d["roofline_photometric"]["backward"])
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_p.err").read()[-1500:])
PY
