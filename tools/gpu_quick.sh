#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu -x -k "lockstep or golden" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency --dump-ops > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | cut -c1-180
grep -o "\"node/neck\": [0-9.]*\|\"dw_multi/detect\": [0-9.]*\|\"stem/backbone\": [0-9.]*" gpurun_out/bench_quick.log
