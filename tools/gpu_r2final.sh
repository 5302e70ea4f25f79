#!/bin/bash
# final evidence of the round: whole GPU suite, smoke, bench line with all legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
grep -E "passed|failed|FAILED|ERROR|pytest exit" gpurun_out/pytest_gpu_final.log | tail -10 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"],2), " conv frac", round(d["roofline"]["frac"],4), "per-kernel", round(d["roofline"]["frac_of_per_kernel_bound"],4), "photo frac", round(d["roofline_photometric"]["frac"],4), "bwd", round(d["roofline_photometric"]["backward"]["frac"],4))
    print("cpu", d.get("cpu_baseline",{}).get("value"), "eager", d.get("gpu_eager_baseline",{}).get("value"), "launches", d["gpu_launches"])
    for k, v in d["kernels"].items(): print("  %-18s %8.3f ms/step  n=%4d  %8.2f us" % (k, v["ms_per_step"], v["launches_per_step"], v["ms_per_launch"]*1e3))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/bench_final.err").read()[-1500:])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null; tail -c 400 gpurun_out/bench_ref_final.json
