#!/bin/bash
# 8 GPUs of one box: training step (graph + coalesced all-reduce; eager + bucketed overlap) and the inference line (e2e through uint8 frames)
N=${1:-8}
mkdir -p gpurun_out
tools/gpu_train_multi.sh $N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E "^\{" | tail -1 | tee gpurun_out/bench_infer_n$N.log | cut -c1-300
