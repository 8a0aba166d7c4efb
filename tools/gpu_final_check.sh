#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"
grep '^{' gpurun_out/bench_n1.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"
grep '^{' gpurun_out/bench_n2.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"
grep '^{' gpurun_out/bench_ref_n2.log | cut -c1-200
