#!/bin/bash
# 2-GPU call: NCCL gradient-exchange test, bench at N=2 with / without the overlapped bucketed all-reduce, C5 at N=1 and N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -k "nccl" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
grep -E "passed|failed|FAILED|ERROR|skipped|pytest exit|^E  " gpurun_out/pytest_nccl.log | tail -8
for ov in 1; do
  JPB_OVERLAP_ALLREDUCE=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_ov$ov.json 2> gpurun_out/bench_n2_ov$ov.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n2_ov$ov.json").read().strip().splitlines()[-1])
    print("N=2 overlap=$ov", round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2))
except Exception as e:
    print("bench N=2 ov=$ov unreadable", e); print(open("gpurun_out/bench_n2_ov$ov.err").read()[-2500:])
PY
done
tail -3 gpurun_out/bench_n2_ov1.err
