#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --mode train --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tail -2 | tee gpurun_out/bench_train_n8_overlap.log | cut -c1-200; echo "exit ${PIPESTATUS[0]}"
