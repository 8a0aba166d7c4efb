#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/conv_probe.py > gpurun_out/conv_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/nms_probe.py > gpurun_out/nms_probe.log 2>&1; echo "nmsprobe rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 3 --dump-ops > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -12 gpurun_out/t_gpu.log; cat gpurun_out/conv_probe.txt; cat gpurun_out/nms_probe.log | tail -12; tail -2 gpurun_out/bench.log
