#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
for mode in 0 1; do
  HN_PDL=$mode timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_pdl$mode.log 2>&1; echo "bench pdl=$mode rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
tail -3 gpurun_out/t_gpu.log
for mode in 0 1; do tail -1 gpurun_out/bench_pdl$mode.log | cut -c1-200; grep -o '"latency_b1_ms": {[^}]*}' gpurun_out/bench_pdl$mode.log; done
