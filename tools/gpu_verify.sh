#!/bin/bash
# what the driver runs at round end: smoke, the GPU test suite, the default bench line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?"; tail -2 gpurun_out/t_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench.log | tail -1
