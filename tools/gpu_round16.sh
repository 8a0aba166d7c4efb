#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 3 --dump-ops > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_nms2_build_kernel -c 1 -f -o gpurun_out/prof_nms_build python tools/profile_step.py 32 > gpurun_out/ncu_full_build.log 2>&1; echo "ncu-build rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_nms2_rounds_kernel -c 1 -f -o gpurun_out/prof_nms_rounds python tools/profile_step.py 32 > gpurun_out/ncu_full_rounds.log 2>&1; echo "ncu-rounds rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/nms_probe.py > gpurun_out/nms_probe.log 2>&1; echo "nmsprobe rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/t_gpu.log; tail -1 gpurun_out/bench.log | cut -c1-300; tail -12 gpurun_out/nms_probe.log
