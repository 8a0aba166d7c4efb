#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_n8.log 2>&1; echo "bench n8 rc=$?"
grep '^{' gpurun_out/bench_n8.log | cut -c1-260
