"""ORACLE -- test infrastructure only.  The reference's CPU path for the hot path, restated with the
reference's own third-party calls (PyTorch fp32 CPU ops, torchvision ``batched_nms``), timed by
``bench.py`` as the ``cpu_baseline`` / ``--impl reference`` arm.  The live reference cannot travel to
the GPU box, so ``kind`` is "port".

Recipe (demo.py:191-244): forward -> lane decode (0.90 / 80) -> seg argmax -> det decode (0.4 / 0.3).
"""
import time

import numpy as np
import torch

from . import hydranet_ref, postproc_ref


def det_decode_torch(x_hw, regression, classification, anchors, threshold, iou_threshold):
    """postprocess() (detection_loss.py:70-108) with torchvision's batched_nms, as the reference calls it."""
    from torchvision.ops.boxes import batched_nms
    a = anchors.expand(regression.shape[0], -1, -1)
    yca, xca = (a[..., 0] + a[..., 2]) / 2, (a[..., 1] + a[..., 3]) / 2
    ha, wa = a[..., 2] - a[..., 0], a[..., 3] - a[..., 1]
    w, h = regression[..., 3].exp() * wa, regression[..., 2].exp() * ha
    yc, xc = regression[..., 0] * ha + yca, regression[..., 1] * wa + xca
    boxes = torch.stack([xc - w / 2., yc - h / 2., xc + w / 2., yc + h / 2.], dim=2)
    boxes[:, :, 0].clamp_(min=0); boxes[:, :, 1].clamp_(min=0)
    boxes[:, :, 2].clamp_(max=x_hw[1] - 1); boxes[:, :, 3].clamp_(max=x_hw[0] - 1)
    scores = torch.max(classification, dim=2, keepdim=True)[0]
    over = (scores > threshold)[:, :, 0]
    out = []
    for i in range(regression.shape[0]):
        if over[i].sum() == 0:
            out.append({'rois': np.array(()), 'class_ids': np.array(()), 'scores': np.array(())})
            continue
        cp = classification[i, over[i, :], ...].permute(1, 0)
        bp, sp = boxes[i, over[i, :], ...], scores[i, over[i, :], ...]
        s_, c_ = cp.max(dim=0)
        keep = batched_nms(bp, sp[:, 0], c_, iou_threshold=iou_threshold)
        out.append({'rois': bp[keep].numpy(), 'class_ids': c_[keep].numpy(), 'scores': s_[keep].numpy()})
    return out


def run_once(sd, cfg, x, det_thr=(0.4, 0.3), lane_thr=(0.90, 80)):
    """One pass of the hot path over a batch on the CPU; returns a small dict of results."""
    with torch.no_grad():
        out = hydranet_ref.forward(sd, cfg, x)
        H, W = x.shape[2:]
        seg = torch.argmax(out["seg"], dim=1).numpy()
        det = det_decode_torch((H, W), out["detection"]["regression"], out["detection"]["classification"],
                               out["detection"]["anchors"], *det_thr)
        lc = cfg["lane"]
        ppl = int(H / lc["interval"])
        fh, fw = int(H / lc["anchor_stride"]), int(W / lc["anchor_stride"])
        lanes = []
        for b in range(x.shape[0]):
            prob = torch.softmax(out["lane"]["predict_cls"][b], -1).numpy()
            lanes.append(postproc_ref.lane_decode_nms(prob, out["lane"]["predict_loc"][b].numpy(), fh, fw, ppl, lc["anchor_stride"],
                                                      float(H) / ppl, W, H, lane_thr[0], lane_thr[1], False, cls_is_prob=True))
    return {"seg": seg, "det": det, "lanes": lanes}


def time_baseline(sd, cfg, batch, steps, warmup, H=640, W=640, seed=0):
    """images/s of the CPU path over ``steps`` batches of ``batch`` images (after ``warmup``)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, H, W, generator=g)
    for _ in range(warmup):
        run_once(sd, cfg, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        run_once(sd, cfg, x)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt
