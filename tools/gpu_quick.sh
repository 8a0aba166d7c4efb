#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_postproc.py -q -m gpu -x 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | cut -c1-180
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_quick2.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick2.log | cut -c1-180
