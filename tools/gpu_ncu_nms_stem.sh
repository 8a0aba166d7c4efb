#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_nms2_build_kernel -c 1 -f -o gpurun_out/prof_nms_build2 python tools/profile_step.py 32 > gpurun_out/ncu_full_build.log 2>&1; echo "ncu-build rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hn_stem_kernel -c 1 -f -o gpurun_out/prof_stem python tools/profile_step.py 32 > gpurun_out/ncu_full_stem.log 2>&1; echo "ncu-stem rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
