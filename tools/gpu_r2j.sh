#!/bin/bash
# 8-GPU call: C2 with / without the overlapped all-reduce, C3 at batch 8/GPU, C5 strong scaling (3/GPU)
mkdir -p gpurun_out
run() { # name, env, args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 $3 > gpurun_out/bench_n8_$1.json 2> gpurun_out/bench_n8_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n8_$1.json").read().strip().splitlines()[-1])
    print("N=8 $1", d["config"]["workload"][:70], round(d["value"], 2), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2))
except Exception as e:
    print("bench N=8 $1 unreadable", e); print(open("gpurun_out/bench_n8_$1.err").read()[-1500:])
PY
}
run C2_overlap JPB_OVERLAP_ALLREDUCE=1 ""
run C2_single JPB_OVERLAP_ALLREDUCE=0 ""
run C3_single JPB_OVERLAP_ALLREDUCE=0 "--config C3"
run C3_overlap JPB_OVERLAP_ALLREDUCE=1 "--config C3"
run C5_single JPB_OVERLAP_ALLREDUCE=0 "--config C5"
