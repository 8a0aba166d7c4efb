"""losses.py (the batched, sync-free restatement of model.py:201-264's loss modules) against the LIVE reference losses:
values pinned here were produced by head_detect/detection_loss.py:FocalLoss, head_seg/segmentation_loss.py:CrossEntropyLoss and
head_lane/lanedetect_loss.py in this container (torch CPU RNG streams are reproducible, so the inputs regenerate anywhere);
when /root/reference is present the comparison (values AND gradients) is repeated against the live code."""
import sys

import pytest
import torch

import hydranet_b200 as hb
from hydranet_b200 import losses
from oracle import ref_live, train_golden

W5 = [0.1, 0.5, 1.0, 5.0, 5.0]
PINNED = {"det_cls": 66036.0703125, "det_reg": 0.16130079329013824, "seg_topk": 11.746953964233398, "seg_focal": 3.5203983783721924,
          "seg_plain": 4.582637786865234, "lane_pos": 9.30213451385498, "lane_neg": 329.1001281738281, "lane_loc": 1.3259562253952026}


def _inputs():
    torch.manual_seed(0)
    B = 3
    gt = train_golden.synthetic_gt(B, 640, 640, 20, 20, 80, seed=5)
    gt["gt_det"][2, :, 4] = -1  # an image without boxes
    cls = torch.sigmoid(torch.randn(B, 76725, 9)).requires_grad_()
    reg = (0.3 * torch.randn(B, 76725, 4)).requires_grad_()
    anc = torch.from_numpy(hb.make_anchors((640, 640), 2.0, [8, 16, 32, 64, 128], [2 ** 0, 2 ** 0.333, 2 ** 0.667], [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)]))
    seg = torch.randn(B, 5, 64, 64).requires_grad_()
    tg = torch.randint(0, 5, (B, 64, 64))
    pc = torch.randn(B, 400, 2).requires_grad_()
    pl = torch.randn(B, 400, 162).requires_grad_()
    return gt, cls, reg, anc, seg, tg, pc, pl


def _ours(gt, cls, reg, anc, seg, tg, pc, pl):
    c, r = losses.detection_loss(cls, reg, anc, gt["gt_det"])
    out = {"det_cls": c.mean(), "det_reg": r.mean()}
    for name, topk, focal in (("seg_topk", True, False), ("seg_focal", False, True), ("seg_plain", False, False)):
        out[name] = losses.seg_loss(seg, tg, torch.tensor(W5), topk, 0.3, focal)
    pos, neg, pm, pn = losses.lane_cls_loss(gt["gt_cls"], pc)
    out.update(lane_pos=pos, lane_neg=neg, lane_loc=losses.lane_reg_loss(pm, pn, gt["gt_loc"], pl))
    return out


def test_losses_match_pinned_reference_values():
    out = _ours(*_inputs())
    for k, v in PINNED.items():
        assert abs(float(out[k]) - v) <= 2e-6 * abs(v), (k, float(out[k]), v)


@pytest.mark.skipif(not ref_live.available(), reason="/root/reference not present")
def test_losses_match_live_reference_values_and_gradients():
    ref_live.import_reference()
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        sys.path.insert(0, "/root/reference/model")
        from head_detect.detection_loss import FocalLoss
        from head_lane.lanedetect_loss import cal_loss_cls, cal_loss_regress
        from head_seg.segmentation_loss import CrossEntropyLoss
        args = _inputs()
        gt, cls, reg, anc, seg, tg, pc, pl = args
        ours = _ours(*args)
        c0, r0 = FocalLoss()(cls, reg, anc, gt["gt_det"])
        ref = {"det_cls": c0.mean(), "det_reg": r0.mean()}
        for name, topk, focal in (("seg_topk", True, False), ("seg_focal", False, True), ("seg_plain", False, False)):
            ref[name] = CrossEntropyLoss(torch.tensor(W5), use_top_k=topk, top_k_ratio=0.3, use_focal=focal)(seg, tg)
        a0, b0, pm, pn = cal_loss_cls(gt["gt_cls"], pc)
        ref.update(lane_pos=a0, lane_neg=b0, lane_loc=cal_loss_regress(pm, pn, gt["gt_loc"], pl))
        leaves = [cls, reg, seg, pc, pl]
        for k in PINNED:
            assert abs(float(ours[k]) - float(ref[k])) <= 2e-6 * abs(float(ref[k])), k
            ga = torch.autograd.grad(ours[k], leaves, retain_graph=True, allow_unused=True)
            gb = torch.autograd.grad(ref[k], leaves, retain_graph=True, allow_unused=True)
            for x, y in zip(ga, gb):
                assert (x is None) == (y is None), k
                if x is not None:
                    assert float((x - y).abs().max()) <= 1e-6 * max(1.0, float(y.abs().max())), k
    finally:
        torch.Tensor.cuda = saved
