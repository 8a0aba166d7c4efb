#!/bin/bash
# A/B of tap runs (A boxes shared by dy taps): off / default policy / everywhere
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
HN_TAP_RUNS=2 timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -x > gpurun_out/t_gpu_runs2.log 2>&1; echo "pytest-gpu-runs2 rc=$?" >> gpurun_out/summary.txt
for mode in 0 1 2; do
  HN_TAP_RUNS=$mode timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-latency --dump-ops > gpurun_out/bench_runs$mode.log 2>&1; echo "bench runs=$mode rc=$?" >> gpurun_out/summary.txt
  cp gpurun_out/op_times.txt gpurun_out/op_times_runs$mode.txt
done
cat gpurun_out/summary.txt
tail -3 gpurun_out/t_gpu.log; tail -3 gpurun_out/t_gpu_runs2.log
for mode in 0 1 2; do tail -1 gpurun_out/bench_runs$mode.log | cut -c1-200; done
