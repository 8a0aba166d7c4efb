#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
tail -4 gpurun_out/t_gpu.log
for m in "--no-fuse-postproc" ""; do
  timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 $m > gpurun_out/bench_fuse_${m:5:4}.log 2>&1; echo "bench [$m] rc=$?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/bench_fuse_${m:5:4}.log | cut -c1-180; grep -o '"latency_b1_ms": {[^}]*}' gpurun_out/bench_fuse_${m:5:4}.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_fuse_${m:5:4}.log
done
cat gpurun_out/summary.txt
