#!/bin/bash
# N-GPU check of the final code: inference line (batch-sharded) and training line (graph with the captured bucketed all-reduce)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E "^\{|Error|error" | tail -2 | tee gpurun_out/bench_infer_n$N.json | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --mode train --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tail -2 | tee gpurun_out/bench_train_n${N}_final.json | cut -c1-300
