"""Kernel-time breakdown of the training step (torch.profiler / CUPTI): python tools/profile_train.py [--batch 16] [--out file]"""
import argparse
import os
import sys

os.environ.setdefault("HN_PROFILE_NAMES", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import hydranet_b200 as hb
from hydranet_b200.config import big_cfg
from oracle import train_golden


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_profile.txt"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = big_cfg()
    torch.manual_seed(0)
    m = hb.HydraNet(cfg).to(dev).train()
    opt = hb.FusedAdam(m.parameters(), lr=1e-5, weight_decay=1e-8)
    B = args.batch
    x = torch.randn(B, 3, 640, 640, device=dev)
    gt = {k: v.to(dev) for k, v in train_golden.synthetic_gt(B, 640, 640, 20, 20, 80, seed=5).items()}

    def step():
        out = m(x)
        ld = m.cal_loss(out, gt)
        loss = train_golden.total_loss(cfg, ld)
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type is not None]
    rows = sorted(((e.device_time_total / 2e3, e.count // 2, e.key) for e in prof.key_averages() if getattr(e, "self_device_time_total", 0) > 0),
                  key=lambda r: -r[0])
    # kernels only: self device time
    rows = sorted(((e.self_device_time_total / 2e3, e.count // 2, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), key=lambda r: -r[0])
    tot = sum(r[0] for r in rows)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("training step, batch %d: %.2f ms of device time per step in %d launches\n" % (B, tot, sum(r[1] for r in rows)))
        for ms, n, k in rows[:60]:
            f.write("%9.3f ms %6d x  %s\n" % (ms, n, k[:110]))
        f.write("\nweight-gradient kernel per layer (device time of the launches under each record_function range)\n")
        wg = sorted(((e.device_time_total / 2e3, e.key) for e in prof.key_averages() if e.key.startswith("wgrad:")), key=lambda r: -r[0])
        for ms, k in wg[:40]:
            f.write("%9.3f ms  %s\n" % (ms, k))
    print(open(args.out).read())


if __name__ == "__main__":
    main()
