"""Training objectives of the three heads (reference: model/model.py:201-264 ``cal_loss`` and the loss modules it calls).

The loss arithmetic stays PyTorch autograd on the heads' fp32 output tensors; what changes
against the reference is the formulation: everything is batched and free of host synchronisation -- no per-image Python
loop, no boolean-mask indexing, no ``int(tensor)`` -- so that a training step can be enqueued (or captured) without the
host waiting for the device.  Semantics restated, with the reference's quirks kept:

  segmentation   head_seg/segmentation_loss.py:27-65   weighted CE (ignore_index 255) -> mean of the top-k hardest pixels
                                                        per image, or the focal variant (small cfg)
  detection      head_detect/detection_loss.py:111-267  focal loss (alpha .25, gamma 2) on IoU-assigned anchors (< 0.4
                                                        negative, >= 0.5 positive, in between ignored) + smooth-L1 (beta 1/9)
                                                        on (dy, dx, dh, dw); per image, then averaged
  lane           head_lane/lanedetect_loss.py:18-78     OHEM cross-entropy (15 negatives per positive, x10) + Huber on the
                                                        positive anchors with ``points_per_line = 160`` hard-coded
                                                        (lanedetect_loss.py:57: the end-position columns that get the x10
                                                        weight are 160 / 161 whatever the configured resolution)
The reference calls ``exit()`` when a loss is zero or not finite (model.py:212-258); that check needs the value on the
host, so it is opt-in here (``model.check_loss_finite = True``) and raises instead of killing the process.
"""
import torch
import torch.nn.functional as F


def seg_loss(prediction, target, class_weights, use_top_k, top_k_ratio, use_focal, ignore_index=255, gamma=2.0, alpha=1.0):
    b = prediction.shape[0]
    w = class_weights.to(prediction.device, prediction.dtype)
    if use_focal:
        eps = 1e-8
        soft = F.softmax(prediction, dim=1) + eps
        one_hot = torch.zeros_like(prediction).scatter_(1, target.unsqueeze(1), 1.0) + eps
        focal = -alpha * torch.pow(1.0 - soft, gamma) * torch.log(soft) * w.view(1, -1, 1, 1)
        loss = torch.sum(one_hot * focal, dim=1).view(b, -1)
    else:
        loss = F.cross_entropy(prediction, target, ignore_index=ignore_index, reduction="none", weight=w).view(b, -1)
        if use_top_k:
            k = int(top_k_ratio * loss.shape[1])
            loss = torch.topk(loss, k, dim=1, sorted=False).values  # == sort descending, keep the first k
    return torch.mean(loss)


class NativeSegLoss(torch.autograd.Function):
    """``seg_loss`` with ``use_top_k`` as a native op (``hn_seg_loss_fwd/bwd``): cross-entropy kernel, 3-pass radix select of the
    k-th largest value per image (no sort), deterministic sums; one backward kernel writes the logits' gradient.  Ties at the
    threshold share their weight (torch.topk keeps an unspecified subset: same loss value, an equally valid subgradient)."""

    @staticmethod
    def forward(ctx, prediction, target, class_weights, top_k_ratio, ignore_index=255):
        from . import _native as nv
        x = prediction.detach().float().contiguous()
        N, Cc, H, W = x.shape
        dev = x.device
        t = target.detach().to(dev, torch.int64).contiguous()
        w = class_weights.detach().to(dev, torch.float32).contiguous()
        k = int(top_k_ratio * (H * W))
        ws = torch.empty(nv.lib.hn_seg_loss_workspace_bytes(N, H * W), dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        d = nv.SegLossDesc(x.data_ptr(), t.data_ptr(), w.data_ptr(), N, Cc, H * W, k, int(ignore_index), ws.data_ptr(), ws.numel(),
                           loss.data_ptr(), None, None)
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_seg_loss_fwd(d, torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(x, t, w, ws)
        ctx.k, ctx.ignore = k, int(ignore_index)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        from . import _native as nv
        x, t, w, ws = ctx.saved_tensors
        N, Cc, H, W = x.shape
        dev = x.device
        gout = g.detach().to(dev, torch.float32).reshape(1).contiguous()
        dx = torch.empty_like(x)
        d = nv.SegLossDesc(x.data_ptr(), t.data_ptr(), w.data_ptr(), N, Cc, H * W, ctx.k, ctx.ignore, ws.data_ptr(), ws.numel(),
                           None, gout.data_ptr(), dx.data_ptr())
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_seg_loss_bwd(d, torch.cuda.current_stream(dev).cuda_stream))
        return dx, None, None, None, None


def detection_loss(classifications, regressions, anchors, annotations, alpha=0.25, gamma=2.0):
    """annotations: [B, M, 5] (x1, y1, x2, y2, class), rows with class -1 are padding (dataloader.py:593-609)."""
    dtype = classifications.dtype
    anchor = anchors[0].to(dtype)                                    # [A, 4] (y1, x1, y2, x2)
    aw, ah = anchor[:, 3] - anchor[:, 1], anchor[:, 2] - anchor[:, 0]
    acx, acy = anchor[:, 1] + 0.5 * aw, anchor[:, 0] + 0.5 * ah
    ann = annotations.to(classifications.device, dtype)
    valid = ann[:, :, 4] != -1                                        # [B, M]
    has_box = valid.any(dim=1)                                        # [B]
    cls = torch.clamp(classifications, 1e-4, 1.0 - 1e-4)             # [B, A, K]
    # IoU of every anchor with every (valid) box
    area = (ann[:, :, 2] - ann[:, :, 0]) * (ann[:, :, 3] - ann[:, :, 1])                                   # [B, M]
    iw = torch.min(anchor[None, :, None, 3], ann[:, None, :, 2]) - torch.max(anchor[None, :, None, 1], ann[:, None, :, 0])
    ih = torch.min(anchor[None, :, None, 2], ann[:, None, :, 3]) - torch.max(anchor[None, :, None, 0], ann[:, None, :, 1])
    inter = torch.clamp(iw, min=0) * torch.clamp(ih, min=0)                                                # [B, A, M]
    ua = torch.clamp((aw * ah)[None, :, None] + area[:, None, :] - inter, min=1e-8)
    iou = torch.where(valid[:, None, :], inter / ua, torch.full_like(inter, -1.0))
    iou_max, iou_arg = torch.max(iou, dim=2)                                                               # [B, A]
    positive = iou_max >= 0.5
    num_pos = positive.sum(dim=1).to(dtype)                                                                # [B]
    assigned = torch.gather(ann, 1, iou_arg[:, :, None].expand(-1, -1, 5))                                 # [B, A, 5]
    # classification targets: -1 ignore, 0 negative, one-hot positive
    targets = torch.full_like(cls, -1.0)
    targets = torch.where((iou_max < 0.4)[:, :, None], torch.zeros_like(targets), targets)
    one_hot = F.one_hot(assigned[:, :, 4].clamp(min=0).long(), cls.shape[2]).to(dtype)
    targets = torch.where(positive[:, :, None], one_hot, targets)
    alpha_factor = torch.where(targets == 1.0, torch.full_like(cls, alpha), torch.full_like(cls, 1.0 - alpha))
    focal_weight = alpha_factor * torch.pow(torch.where(targets == 1.0, 1.0 - cls, cls), gamma)
    bce = -(targets * torch.log(cls) + (1.0 - targets) * torch.log(1.0 - cls))
    cls_loss = torch.where(targets != -1.0, focal_weight * bce, torch.zeros_like(cls)).sum(dim=(1, 2)) / torch.clamp(num_pos, min=1.0)
    # images without boxes: every anchor is a negative, NOT normalised (detection_loss.py:138-160)
    empty_loss = ((1.0 - alpha) * torch.pow(cls, gamma) * -torch.log(1.0 - cls)).sum(dim=(1, 2))
    cls_loss = torch.where(has_box, cls_loss, empty_loss)
    # box regression on the positive anchors
    gw, gh = assigned[:, :, 2] - assigned[:, :, 0], assigned[:, :, 3] - assigned[:, :, 1]
    gcx, gcy = assigned[:, :, 0] + 0.5 * gw, assigned[:, :, 1] + 0.5 * gh
    gw, gh = torch.clamp(gw, min=1), torch.clamp(gh, min=1)
    t = torch.stack(((gcy - acy) / ah, (gcx - acx) / aw, torch.log(gh / ah), torch.log(gw / aw)), dim=2)  # (dy, dx, dh, dw)
    diff = torch.abs(t - regressions)
    sl1 = torch.where(diff <= 1.0 / 9.0, 0.5 * 9.0 * diff * diff, diff - 0.5 / 9.0)
    sl1 = torch.where(positive[:, :, None], sl1, torch.zeros_like(sl1)).sum(dim=(1, 2))
    reg_loss = torch.where(positive.any(dim=1) & has_box, sl1 / torch.clamp(4.0 * num_pos, min=1.0), torch.zeros_like(sl1))
    return cls_loss.mean(dim=0, keepdim=True), reg_loss.mean(dim=0, keepdim=True)


class NativeDetectionLoss(torch.autograd.Function):
    """The detection objective as ONE native op (``hn_det_loss``: assignment, focal + smooth-L1 terms and their gradients in
    three launches instead of ~60 torch kernels over [B, 76 725, M] intermediates).  Returns (mean_b cls_loss_b, mean_b reg_loss_b),
    the two scalars ``cal_loss`` puts into the loss dict (model.py:224-235).  Checked against ``detection_loss`` above (which is
    pinned against the live reference) in tests/test_gpu_train_ops.py."""

    @staticmethod
    def forward(ctx, classifications, regressions, anchors, annotations, alpha=0.25, gamma=2.0):
        import ctypes as C

        from . import _native as nv
        cls = classifications.detach().float().contiguous()
        reg = regressions.detach().float().contiguous()
        B, A, K = cls.shape
        dev = cls.device
        anc = anchors.detach().to(dev, torch.float32).reshape(-1, 4).contiguous()
        ann = annotations.detach().to(dev, torch.float32).contiguous()
        M = ann.shape[1]
        ws = torch.empty(B * A * 4 + B * 24 + 64, dtype=torch.uint8, device=dev)
        out = torch.empty((2, B), dtype=torch.float32, device=dev)
        dcls, dreg = torch.empty_like(cls), torch.empty_like(reg)
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_det_loss(cls.data_ptr(), reg.data_ptr(), anc.data_ptr(), ann.data_ptr(), B, A, K, M, float(alpha), float(gamma),
                                        ws.data_ptr(), ws.numel(), out[0].data_ptr(), out[1].data_ptr(), dcls.data_ptr(), dreg.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(dcls, dreg)
        m = out.mean(dim=1)
        return m[0], m[1]

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        dcls, dreg = ctx.saved_tensors
        return dcls * g_cls, dreg * g_reg, None, None, None, None


def lane_cls_loss(cls_targets, cls_preds, negative_ratio=15, alpha=10):
    t = cls_targets[..., 1].reshape(-1)
    pmask = t > 0
    nmask = ~pmask
    fp, fn = pmask.float(), nmask.float()
    preds = cls_preds.reshape(-1, cls_preds.shape[-1])
    n_pos, n_neg = fp.sum(), fn.sum()
    negative_num = torch.minimum(torch.clamp(n_pos * negative_ratio, min=1), n_neg).long()  # clamp(x, min=1, max=n_neg)
    positive_num = torch.clamp(n_pos, min=1).int()
    logp = F.log_softmax(preds, dim=-1)
    fg, bg = logp[..., 1], logp[..., 0]
    # k-th smallest background log-probability among the negatives (the hardest `negative_num` negatives), without a sync:
    # positives are pushed to +inf, the index is a device tensor
    ranked = torch.sort(torch.where(nmask, bg.detach(), torch.full_like(bg, float("inf")))).values
    kth = ranked.index_select(0, torch.clamp(negative_num - 1, max=ranked.numel() - 1).reshape(1)).squeeze(0)  # (ranked[tensor] would sync)
    ohem = (bg <= kth).float() * fn
    total_pos = -torch.sum(alpha * fg * fp) / positive_num
    total_neg = -torch.sum(alpha * bg * ohem) / positive_num
    return total_pos, total_neg, pmask, positive_num


def lane_reg_loss(pmask, positive_num, loc_targets, loc_preds, alpha=10, points_per_line=160):
    preds = loc_preds.reshape(-1, loc_preds.shape[-1])
    tgt = loc_targets.reshape(-1, loc_targets.shape[-1])
    weight = torch.ones_like(tgt)
    weight[..., points_per_line + 1] = alpha
    weight[..., points_per_line] = alpha
    valid_pts = (tgt != 0).float()
    mask = weight * pmask.unsqueeze(-1).float() * valid_pts
    d = preds - tgt
    ad = torch.abs(d)
    huber = torch.where(ad < 1, d * d / 2, ad - 0.5) * mask
    per_anchor = torch.sum(huber, -1) / torch.sum(valid_pts, -1).clamp(min=1)
    return torch.sum(per_anchor) / positive_num


class NativeLaneLoss(torch.autograd.Function):
    """``lane_cls_loss`` + ``lane_reg_loss`` as one native op (``hn_lane_loss``, three launches instead of ~40 torch kernels and a
    sort): returns (total_pos, total_neg, loc)."""

    @staticmethod
    def forward(ctx, cls_targets, cls_preds, loc_targets, loc_preds, negative_ratio=15, alpha=10, points_per_line=160):
        from . import _native as nv
        dev = cls_preds.device
        cp = cls_preds.detach().float().reshape(-1, 2).contiguous()
        ct = cls_targets.detach().to(dev, torch.float32).reshape(-1, 2).contiguous()
        L = loc_preds.shape[-1]
        lp = loc_preds.detach().float().reshape(-1, L).contiguous()
        lt = loc_targets.detach().to(dev, torch.float32).reshape(-1, L).contiguous()
        T = cp.shape[0]
        ws = torch.empty(nv.lib.hn_lane_loss_workspace_bytes(T), dtype=torch.uint8, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        dcls = torch.empty((2, T, 2), dtype=torch.float32, device=dev)
        dloc = torch.empty((T, L), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_lane_loss(ct.data_ptr(), cp.data_ptr(), lt.data_ptr(), lp.data_ptr(), T, L, int(points_per_line), float(negative_ratio),
                                         float(alpha), ws.data_ptr(), ws.numel(), out.data_ptr(), dcls.data_ptr(), dloc.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(dcls, dloc)
        ctx.shapes = (cls_preds.shape, loc_preds.shape)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_pos, g_neg, g_loc):
        dcls, dloc = ctx.saved_tensors
        cs, ls = ctx.shapes
        return None, (dcls[0] * g_pos + dcls[1] * g_neg).reshape(cs), None, (dloc * g_loc).reshape(ls), None, None, None


def cal_loss(model, pred_dict, gt_dict):
    """model/model.py:201-264: dict of the heads' loss terms (the caller weights and sums them, train.py:192-203)."""
    out = {}
    if model.train_seg:
        sc = model.cfgs["segment"]
        if sc.get("use_lovasz", False):
            raise NotImplementedError("use_lovasz=True is not configured by any reference cfg (model/cfgs/*.yml)")
        dev = pred_dict["seg"].device
        cache = model.__dict__.setdefault("_loss_consts", {})
        if ("seg_w", dev) not in cache:  # uploaded once: a CUDA-graph capture of the step must not see host->device copies
            cache[("seg_w", dev)] = torch.tensor(sc["class_weight"], dtype=torch.float32, device=dev)
        if (pred_dict["seg"].is_cuda and getattr(model, "native_seg_loss", True) and sc["use_top_k"] and not sc["use_focal"]
                and int(sc["top_k_ratio"] * pred_dict["seg"].shape[2] * pred_dict["seg"].shape[3]) >= 1):
            out["loss_seg"] = NativeSegLoss.apply(pred_dict["seg"], gt_dict["gt_seg"], cache[("seg_w", dev)], sc["top_k_ratio"])
        else:
            out["loss_seg"] = seg_loss(pred_dict["seg"], gt_dict["gt_seg"].to(dev).long(), cache[("seg_w", dev)],
                                       sc["use_top_k"], sc["top_k_ratio"], sc["use_focal"])
    if model.train_detect:
        det = pred_dict["detection"]
        if det["classification"].is_cuda and getattr(model, "native_det_loss", True) and gt_dict["gt_det"].shape[1] <= 64:
            out["loss_det_cls"], out["loss_det_reg"] = NativeDetectionLoss.apply(det["classification"], det["regression"], det["anchors"], gt_dict["gt_det"])
        else:
            c, r = detection_loss(det["classification"], det["regression"], det["anchors"], gt_dict["gt_det"])
            out["loss_det_cls"], out["loss_det_reg"] = c.mean(), r.mean()
    if model.train_lane:
        lane = pred_dict["lane"]
        dev = lane["predict_cls"].device
        if (lane["predict_cls"].is_cuda and getattr(model, "native_lane_loss", True) and lane["predict_cls"].shape[-1] == 2
                and lane["predict_loc"].shape[-1] >= 162):
            out["loss_lane_cls_pos"], out["loss_lane_cls_neg"], out["loss_lane_loc"] = NativeLaneLoss.apply(
                gt_dict["gt_cls"], lane["predict_cls"], gt_dict["gt_loc"], lane["predict_loc"])
        else:
            pos, neg, pmask, n_pos = lane_cls_loss(gt_dict["gt_cls"].to(dev), lane["predict_cls"])
            out["loss_lane_cls_pos"], out["loss_lane_cls_neg"] = pos, neg
            out["loss_lane_loc"] = lane_reg_loss(pmask, n_pos, gt_dict["gt_loc"].to(dev), lane["predict_loc"])
    if getattr(model, "check_loss_finite", False):
        for k, v in out.items():
            if float(v) == 0 or not torch.isfinite(v):
                raise FloatingPointError("cal %s diverge (model.py:212-258)" % k)
    return out
