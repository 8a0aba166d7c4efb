"""One small eager training step (big cfg, 128x128, batch 2: every training kernel, every conv flavour) + Adam, for compute-sanitizer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hydranet_b200 as hb
from hydranet_b200 import losses
from hydranet_b200.config import big_cfg
from oracle import synth, train_golden

cfg = big_cfg(128, 128)
torch.manual_seed(0)
m = hb.HydraNet(cfg)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1))
m = m.cuda().train()
os.environ["HN_SIDE_WGRAD"] = os.environ.get("HN_SIDE_WGRAD", "1")
opt = hb.FusedAdam(m.parameters(), lr=1e-5, weight_decay=1e-8)
x = synth.synth_input(2, 128, 128, seed=3).cuda()
gt = {k: v.cuda() for k, v in train_golden.synthetic_gt(2, 128, 128, 4, 4, 16, seed=5).items()}
out = m(x)
c, r = losses.NativeDetectionLoss.apply(out["detection"]["classification"], out["detection"]["regression"], out["detection"]["anchors"], gt["gt_det"])
w = torch.tensor(cfg["segment"]["class_weight"], device="cuda")
seg = losses.NativeSegLoss.apply(out["seg"], gt["gt_seg"], w, 0.3)
L = out["lane"]["predict_loc"].shape[-1]
lp, ln, ll = losses.NativeLaneLoss.apply(gt["gt_cls"], out["lane"]["predict_cls"], gt["gt_loc"], out["lane"]["predict_loc"], 15, 10, L - 2)
loss = 5 * seg + c + 50 * r + lp + ln + ll
loss.backward()
opt.step()
torch.cuda.synchronize()
print("sanitize_train ok: loss %.4f" % float(loss))
