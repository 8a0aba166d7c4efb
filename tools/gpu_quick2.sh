#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/bench_quick.log | tail -1 | cut -c1-900; grep -i "error\|Traceback" gpurun_out/bench_quick.log | head -3
