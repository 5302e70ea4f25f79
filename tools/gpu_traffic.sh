#!/bin/bash
# ncu DRAM traffic + duration of the conv / photometric launches of ONE eager single-stream step (the 5th step of the bench
# command: after 3 warm-up steps, inside the 3-step profiling pass).  Parsed by tools/ncu_traffic.py.
mkdir -p gpurun_out
PER=$(python - <<'PY'
print(487 + 8)   # conv_tc launches (176 forward + 164 data gradient + 147 weight gradient) + 8 photometric launches per step
PY
)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc|photometric" \
  -s $((PER * 4)) -c $PER --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-graph > gpurun_out/traffic_bench.log 2>&1
grep -c conv_tc gpurun_out/traffic.csv
