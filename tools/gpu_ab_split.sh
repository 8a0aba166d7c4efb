#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
for mode in 0 1; do
  HN_SPLIT=$mode timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_split$mode.log 2>&1; echo "bench split=$mode rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_gpu.log
for mode in 0 1; do tail -1 gpurun_out/bench_split$mode.log | cut -c1-200; grep -o '"e2e": {[^}]*}' gpurun_out/bench_split$mode.log | tail -1; done
