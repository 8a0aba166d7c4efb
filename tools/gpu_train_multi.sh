#!/bin/bash
# training step on N GPUs (data parallel): graph mode (coalesced all-reduce after the graph) and eager mode (bucketed, overlapped)
N=${1:-2}
mkdir -p gpurun_out
for mode in "" "--no-graph"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --mode train --steps 5 --warmup 3 $mode 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee -a gpurun_out/bench_train_n$N.log
done
