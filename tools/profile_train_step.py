"""One EAGER training step (batch 16) between cudaProfilerStart/Stop, for `ncu --profile-from-start off ...`.
HN_NVTX=1 wraps every weight-gradient launch in an NVTX range "wgrad:<layer>" (ncu --nvtx --nvtx-include)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hydranet_b200 as hb
from hydranet_b200.config import big_cfg
from oracle import train_golden


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda", 0)
    cfg = big_cfg()
    torch.manual_seed(0)
    m = hb.HydraNet(cfg).to(dev).train()
    opt = hb.FusedAdam(m.parameters(), lr=1e-5, weight_decay=1e-8)
    step = hb.TrainStep(m, opt, graph=False)
    x = torch.randn(B, 3, 640, 640, device=dev)
    gt = {k: v.to(dev) for k, v in train_golden.synthetic_gt(B, 640, 640, 20, 20, 80, seed=5).items()}
    for _ in range(2):
        step(x, gt)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(x, gt)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
