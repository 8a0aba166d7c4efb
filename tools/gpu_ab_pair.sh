#!/bin/bash
mkdir -p gpurun_out
for bn in 192 128; do
  HN_PAIR_MIN_BN=$bn timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency --dump-ops > gpurun_out/bench_pair$bn.log 2>&1; echo "bench pair_min_bn=$bn rc=$?"
  tail -1 gpurun_out/bench_pair$bn.log | cut -c1-180
  grep "seg.d4\|seg.d5.p00" gpurun_out/op_times.txt | awk '{print $2,$5}'
done
HN_PAIR_MIN_BN=128 timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -2
