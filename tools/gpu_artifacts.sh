#!/bin/bash
# artefacts for profiles/: tests, canonical bench line, per-op times, ncu launch list (+ DRAM bytes), ncu full of the top conv launches
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest-gpu rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --dump-ops --e2e-probe > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench-ref rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu-list rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_conv_gemm_kernel -s $(awk '$1=="seg.d2"{print $2}' gpurun_out/conv_index.txt) -c 2 -f -o gpurun_out/prof_seg_d2_d3 python tools/profile_step.py 32 > gpurun_out/ncu_full_seg.log 2>&1; echo "ncu-seg rc=$?" >> gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/gpu.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/t_gpu.log; tail -1 gpurun_out/bench.log; tail -1 gpurun_out/bench_ref.log
