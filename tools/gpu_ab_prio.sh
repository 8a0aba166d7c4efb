#!/bin/bash
mkdir -p gpurun_out
for n in 0 1 0 1; do
  HN_BRANCH_PRIO=$n timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_prio$n.log 2>&1; echo "bench prio=$n rc=$?"
  tail -1 gpurun_out/bench_prio$n.log | cut -c1-180
done
