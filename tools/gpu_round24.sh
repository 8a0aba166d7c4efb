#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for mode in "" "--overlap-postproc"; do
  HN_SPLIT=0 timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 --no-latency $mode > gpurun_out/bench_ovl_${mode:2:3}.log 2>&1; echo "bench [$mode] rc=$?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/bench_ovl_${mode:2:3}.log | cut -c1-200
done
cat gpurun_out/summary.txt
