"""ORACLE -- test infrastructure only.  Golden vectors for the TRAINING step,
which the product does not implement yet: this file pins what a future B200 training path has to reproduce.

Runs the LIVE reference (/root/reference, this container only) for one step of train.py:241-269 on CPU:
    net.train(); out = net(x); loss_dict = net.cal_loss(out, gt); total = cal_total_loss (train.py:192-203);
    zero_grad; total.backward(); Adam(lr=1e-5, weight_decay=1e-8).step()          (train.py:147, yml:12-13)
with the deterministic synthetic weights / inputs of oracle/synth.py and synthetic ground truth in the formats the
data loader produces (dataloader.py:593-609 boxes xyxy+class padded with -1; lane_codec.py:221-252 cls one-hot
[400,2] and loc [400,162]); writes tests/golden/train_step_<cfg>.npz:
    loss components and total, global gradient norm, gradient norm per top-level module, gradient and post-step
    digests (first 256 values, sum, L2 norm) of a few probe tensors (first / last layers of every sub-network), BatchNorm
    running statistics after the step.

    python oracle/train_golden.py            # rewrites tests/golden/train_step_*.npz
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_live, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PROBES = ("backbone.net.stem.conv.weight", "backbone.net.stage_4.blocks.block_13.conv_block_3.0.weight",
          "neck.bifpn.0.p4_w2", "neck.bifpn.2.conv3_up.pointwise_conv.conv.weight", "segheader.decoder.8.conv.weight",
          "detectheader.regressor.header.pointwise_conv.conv.bias", "detectheader.classifier.bn_list.4.2.weight",
          "laneheader.conv_up_conv.3.bias")


def synthetic_gt(B, H, W, fh, fw, ppl, seed):
    """Config 4 ground truth: seg labels, padded boxes, lane anchors with a few positives."""
    g = torch.Generator().manual_seed(seed)
    gt_seg = torch.randint(0, 5, (B, H, W), generator=g).float()
    M = 6
    gt_det = -torch.ones((B, M, 5))
    for b in range(B):
        for k in range(3 + b % 2):
            x1, y1 = torch.rand(2, generator=g) * torch.tensor([W * 0.6, H * 0.6])
            w, h = 16 + torch.rand(2, generator=g) * torch.tensor([W * 0.3, H * 0.3])
            gt_det[b, k] = torch.tensor([x1, y1, min(x1 + w, W - 1.0), min(y1 + h, H - 1.0), float(torch.randint(0, 9, (1,), generator=g))])
    na = fh * fw
    gt_cls = torch.zeros((B, na, 2))
    gt_cls[:, :, 0] = 1.0
    gt_loc = torch.zeros((B, na, 2 * ppl + 2))
    for b in range(B):
        pos = torch.randperm(na, generator=g)[:5]
        gt_cls[b, pos, 0], gt_cls[b, pos, 1] = 0.0, 1.0
        gt_loc[b, pos] = torch.randn((5, 2 * ppl + 2), generator=g)
        gt_loc[b, pos, ppl] = torch.randint(1, ppl, (5,), generator=g).float()        # down end position
        gt_loc[b, pos, ppl + 1] = torch.randint(1, ppl, (5,), generator=g).float()    # up end position
    return {"gt_seg": gt_seg, "gt_det": gt_det, "gt_cls": gt_cls, "gt_loc": gt_loc}


def total_loss(cfg, d):
    """train.py:192-203."""
    t = d["loss_seg"] * cfg["segment"]["segment_weight"]
    t = t + (d["loss_det_cls"] * cfg["detection"]["loss_cls_weight"] + d["loss_det_reg"] * cfg["detection"]["loss_reg_weight"]) * cfg["detection"]["detection_weight"]
    t = t + (d["loss_lane_cls_pos"] * cfg["lane"]["loss_cls_pos_weight"] + d["loss_lane_cls_neg"] * cfg["lane"]["loss_cls_neg_weight"]
             + d["loss_lane_loc"] * cfg["lane"]["loss_loc_weight"]) * cfg["lane"]["lane_weight"]
    return t


def run_step(ref_model, cfg, H, W, B=2, seed=1):
    torch.manual_seed(0)
    net = ref_model.HydraNet(cfg)
    net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=seed, seg_logit_gain=1.0))
    net.train()
    x = synth.synth_input(B, H, W, seed=3)
    lh = net.laneheader
    fh, fw = H // cfg["lane"]["anchor_stride"], W // cfg["lane"]["anchor_stride"]
    ppl = int(H / cfg["lane"]["interval"])
    gt = synthetic_gt(B, H, W, fh, fw, ppl, seed=5)
    opt = torch.optim.Adam(net.parameters(), lr=cfg["train"]["lr"], weight_decay=cfg["train"]["weight_decay"])
    out = net(x)
    ld = net.cal_loss(out, gt)
    tot = total_loss(cfg, ld)
    opt.zero_grad()
    tot.backward()
    named = dict(net.named_parameters())
    blob = {"loss_total": np.float64(tot.item())}
    for k, v in ld.items():
        blob[k] = np.float64(v.item())
    gn = {}
    for k, p in named.items():
        if p.grad is None:
            continue
        top = k.split(".")[0]
        gn[top] = gn.get(top, 0.0) + float(p.grad.double().pow(2).sum())
    for top, v in gn.items():
        blob["gradnorm." + top] = np.float64(v ** 0.5)
    blob["gradnorm.all"] = np.float64(sum(gn.values()) ** 0.5)
    blob["n_params_without_grad"] = np.int64(sum(1 for p in named.values() if p.grad is None))
    def digest(prefix, t):  # first 256 values + two sums: a few KB per tensor instead of megabytes
        f = t.detach().double().reshape(-1)
        blob[prefix + ".head"] = f[:256].float().numpy().copy()
        blob[prefix + ".sum"] = np.float64(f.sum())
        blob[prefix + ".l2"] = np.float64(f.pow(2).sum().sqrt())
    for k in PROBES:
        digest("grad." + k, named[k].grad)
    opt.step()
    for k in PROBES:
        digest("after." + k, named[k])
    sd = net.state_dict()
    for k in ("backbone.net.stem.bn.running_mean", "backbone.net.stem.bn.running_var", "neck.bifpn.0.p5_down_channel.1.running_mean"):
        blob["bn." + k] = sd[k].numpy().copy()
    return blob


def main(out_dir=GOLD):
    import yaml
    ref_model, _ = ref_live.import_reference()
    if not torch.cuda.is_available():  # segmentation_loss.py:53 moves its class weights with Tensor.cuda()
        torch.Tensor.cuda = lambda self, *a, **k: self
    torch.set_num_threads(8)
    base = yaml.safe_load(open("/root/reference/model/cfgs/hydranet_joint_big_backbone.yml"))
    cfg = copy.deepcopy(base)
    # 640x640 only: the reference's lane regression loss hard-codes points_per_line = 160 (lanedetect_loss.py:57,65-66),
    # i.e. the 162-column loc tensor of the default input size
    blob = run_step(ref_model, cfg, 640, 640)
    np.savez_compressed(os.path.join(out_dir, "train_step_big_640x640.npz"), **blob)
    print({k: (float(v) if np.ndim(v) == 0 else v.shape) for k, v in blob.items() if not k.startswith(("grad.", "after.", "bn."))})


if __name__ == "__main__":
    main()
