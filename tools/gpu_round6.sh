#!/bin/bash
# 2 GPUs: graph mode with the exchange captured inside the graph vs after it; then the 1-GPU native detection-loss test
for ov in 1 0; do
  HN_GRAPH_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 5 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tail -2 | cut -c1-140
done
timeout 600 python -m pytest tests/test_gpu_train_ops.py -q -x -k "detection_loss or adam" --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --mode train --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-160
