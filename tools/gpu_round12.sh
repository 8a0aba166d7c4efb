#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --e2e-probe > gpurun_out/bench.log 2>&1; echo "bench-default rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench-ref rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log; grep "e2e probe" gpurun_out/bench.log; tail -1 gpurun_out/bench.log; tail -1 gpurun_out/bench_ref.log | cut -c1-300
