#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_conv_gemm_kernel -s 199 -c 1 -f -o gpurun_out/prof_d7 python tools/profile_step.py 32 > gpurun_out/ncu_full_d7.log 2>&1; echo "ncu rc=$?"
