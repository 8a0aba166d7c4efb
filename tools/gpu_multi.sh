#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"
tail -1 gpurun_out/bench_n1.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-seconds 1 --no-latency > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"
tail -1 gpurun_out/bench_n2.log | cut -c1-1200
