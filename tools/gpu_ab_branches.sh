#!/bin/bash
mkdir -p gpurun_out
HN_BRANCHES=1 timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -2
for m in 0 1; do
  HN_BRANCHES=$m timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_br$m.log 2>&1; echo "bench branches=$m rc=$?"
  tail -1 gpurun_out/bench_br$m.log | cut -c1-180; grep -o '"latency_b1_ms": {[^}]*}' gpurun_out/bench_br$m.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_br$m.log
done
