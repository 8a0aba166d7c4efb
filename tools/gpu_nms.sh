#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_postproc.py -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-postproc rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/nms_probe.py > gpurun_out/nms_probe.log 2>&1; echo "nmsprobe rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu-list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_gpu.log; tail -10 gpurun_out/nms_probe.log; grep -E "nms2|det_" gpurun_out/launches.csv | awk -F'","' '{print $5, $(NF)}' | head -20
