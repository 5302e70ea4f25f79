#!/bin/bash
# Round 2, call 2: first device run of the 3xTF32 precision mode and of the un-skipped whole-model cases; per-layer conv table.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_conv.py > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest(a) exit $?" >> gpurun_out/pytest_gpu_a.log
grep -E "passed|failed|FAILED|ERROR|pytest\(a\) exit|TF32 calibration|^  \(|^  [a-z('\"]" gpurun_out/pytest_gpu_a.log | tail -60
timeout 600 python -m pytest tests/test_conv.py -m gpu -q > gpurun_out/pytest_gpu_conv.log 2>&1; echo "pytest(conv) exit $?" >> gpurun_out/pytest_gpu_conv.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu_conv.log | tail -40
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/conv_layers.py > gpurun_out/conv_layers.txt 2>&1; head -45 gpurun_out/conv_layers.txt
