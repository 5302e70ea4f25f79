#!/bin/bash
# Round 2, first GPU call (1 GPU, ~12 min of box time): first on-device run of everything written after the round-1 GPU budget
# was spent, and the A/B of the opt-in kernel variants against the measured defaults.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh'
# Outputs land in gpurun_out/ (copy what is to be judged into profiles/ as r2_*).
mkdir -p gpurun_out
# 1. the whole GPU suite WITHOUT -x, so that one failing first-run case does not hide the others
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|pytest exit" gpurun_out/pytest_gpu.log | tail -15
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
# 2. photometric forward: measured default (2) vs packed fp32 pairs (3), B = 4 and 8
for v in 2 3; do for b in 4 8; do
  timeout 120 python tools/bench_photometric.py --B $b --variant $v > gpurun_out/photo_v${v}_b${b}.json 2> gpurun_out/photo_v${v}_b${b}.err
  python -c "
import json; d=json.load(open('gpurun_out/photo_v${v}_b${b}.json')); print('photometric variant', d['fwd_variant'], 'B', d['B'], 'fwd ms', round(d['fwd_ms'],4), 'frac', round(d['fwd_frac'],4), 'bwd ms', round(d['bwd_ms'],4))"
done; done
# 2b. photometric backward staging the forward's warped frames (opt-in JPB_PHOTO_KEEP_WARPED=1): compare bwd_ms with 2. above
for b in 4 8; do
  JPB_PHOTO_KEEP_WARPED=1 timeout 120 python tools/bench_photometric.py --B $b --variant 2 > gpurun_out/photo_keep_b${b}.json 2> gpurun_out/photo_keep_b${b}.err
  python -c "
import json; d=json.load(open('gpurun_out/photo_keep_b${b}.json')); print('photometric keep-warped B', d['B'], 'fwd ms', round(d['fwd_ms'],4), 'bwd ms', round(d['bwd_ms'],4))"
done
# 2c. prepared patch: backward accumulators in registers (tools/experiments/README.md) — built into a scratch copy of the tree
rm -rf /tmp/jpb_exp && mkdir -p /tmp/jpb_exp && cp -r bench.py tools tests include jperceiver_b200 oracle mono mmcv MEASURED_PEAKS.json /tmp/jpb_exp/ 2>/dev/null
( cd /tmp/jpb_exp && patch -p1 -s < tools/experiments/photometric_bwd_regs.patch && rm -f jperceiver_b200/csrc/libjpb200.so && python -m jperceiver_b200.build > /dev/null 2>&1 \
  && for b in 4 8; do timeout 120 python tools/bench_photometric.py --B $b --variant 2 > $OLDPWD/gpurun_out/photo_bwdregs_b${b}.json 2> $OLDPWD/gpurun_out/photo_bwdregs_b${b}.err; \
     python -c "
import json; d=json.load(open('$OLDPWD/gpurun_out/photo_bwdregs_b${b}.json')); print('photometric bwd-in-registers B', d['B'], 'bwd ms', round(d['bwd_ms'],4))"; done )
# 3. the memory-bound network kernels (BatchNorm, max-pool): achieved GB/s per call at the step's largest shapes
timeout 200 python tools/bench_misc.py > gpurun_out/misc_default.txt 2>&1; tail -12 gpurun_out/misc_default.txt
# 4. the bench line with the default kernels, then with the packed photometric forward
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
JPB_PHOTO_FWD=3 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_photo3.json 2> gpurun_out/bench_n1_photo3.err
JPB_PHOTO_KEEP_WARPED=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_keep.json 2> gpurun_out/bench_n1_keep.err
python - <<'EOF'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_n1_photo3.json", "gpurun_out/bench_n1_keep.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2),
              "conv frac", round(d["roofline"]["frac"], 4), "photo frac", round(d["roofline_photometric"]["frac"], 4))
    except Exception as e:   # noqa: BLE001
        print(f, "unreadable:", e)
EOF
# 5. evidence: ncu --set full of the packed photometric forward (scale 0) and the launch list of one eager step
JPB_PHOTO_FWD=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric_fwd -s 4 -c 1 -o gpurun_out/ncu_photo_v3 -f \
  python tools/bench_photometric.py --B 4 --iters 2 > gpurun_out/ncu_photo_v3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -20
