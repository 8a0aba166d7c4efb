#!/bin/bash
mkdir -p gpurun_out
for n in 3 6 12 1; do
  HN_ROUNDS_PASSES=$n timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency --no-fuse-postproc > gpurun_out/bench_pass$n.log 2>&1; echo "bench passes=$n rc=$?"
  tail -1 gpurun_out/bench_pass$n.log | cut -c1-180
done
HN_ROUNDS_PASSES=6 timeout 900 python -m pytest tests/test_gpu_postproc.py -q -m gpu -x 2>&1 | tail -2
