#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 3 --dump-ops > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu-list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/t_gpu.log; tail -1 gpurun_out/bench.log | cut -c1-300
grep -E "nms2|det_" gpurun_out/launches.csv | awk -F'","' '{print $5, $(NF)}' | head -20
