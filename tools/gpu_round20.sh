#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for spec in "c1:0" "s4c1:110" "detpw:204"; do
  name=${spec%%:*}; skip=${spec##*:}
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_conv_gemm_kernel -s $skip -c 1 -f -o gpurun_out/prof_conv_$name python tools/profile_step.py 32 > gpurun_out/ncu_full_$name.log 2>&1; echo "ncu $name rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
