#!/bin/bash
# first GPU contact: per-op lockstep diagnostics, post-processing parity, forward parity, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -k lockstep > gpurun_out/t_lockstep.log 2>&1; echo "lockstep rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_postproc.py -q > gpurun_out/t_postproc.log 2>&1; echo "postproc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_forward.py -q -k "not lockstep" > gpurun_out/t_forward.log 2>&1; echo "forward rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_lockstep.log; tail -15 gpurun_out/t_postproc.log; tail -15 gpurun_out/t_forward.log; tail -3 gpurun_out/bench.log
