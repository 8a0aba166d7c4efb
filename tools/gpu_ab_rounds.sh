#!/bin/bash
mkdir -p gpurun_out
for n in 2 1; do
  HN_ROUNDS_CTAS=$n timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_rounds$n.log 2>&1; echo "bench rounds_ctas=$n rc=$?"
  tail -1 gpurun_out/bench_rounds$n.log | cut -c1-180
done
