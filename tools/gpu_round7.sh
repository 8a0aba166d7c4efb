#!/bin/bash
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 5 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tail -2 | cut -c1-140; echo "exit ${PIPESTATUS[0]}"
