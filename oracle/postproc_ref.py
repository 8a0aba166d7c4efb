"""ORACLE -- test infrastructure only (never imported by the product path).

numpy restatement of the reference's post-processing, every step in the reference's arithmetic type:
  bbox_transform / clip_boxes   head_detect/detection_loss.py:7-52
  postprocess                   head_detect/detection_loss.py:70-108
  nms / batched_nms             torchvision.ops.boxes (third-party, un-vendored; the image has
                                torchvision 0.26.0+cu128; semantics: stable score-descending order,
                                suppress j if inter/(area_i+area_j-inter) > thr in fp32; coordinate
                                trick when boxes.numel() <= 4000 (CPU) / 100000 (CUDA), else per class)
  decode_lane                   head_lane/lane_codec.py:116-219
  calc_err_dis_with_pos / nms_with_pos   head_lane/lane_codec_utils.py:487-542
  argmax                        head_seg/segmentation.py:109
Pinned against the live reference and torchvision by tests/test_oracle_pinning.py.

One deliberate difference, stated here and in DESIGN.md: ``exp`` in the box decode is evaluated in
float64 and rounded once (correctly rounded fp32); torch's vectorised fp32 exp can differ from that by
1 ulp, so decoded w/h are compared to the live reference with a 2-ulp tolerance and NMS parity is
checked on identical pre-NMS boxes.
"""
import numpy as np

f32 = np.float32


def bbox_transform_clip(anchors, regression, img_h, img_w):
    """anchors [A,4] (y1,x1,y2,x2), regression [N,A,4] (dy,dx,dh,dw) -> boxes [N,A,4] xyxy, clipped."""
    a = anchors.astype(f32)
    r = regression.astype(f32)
    yca = (a[..., 0] + a[..., 2]) / f32(2)
    xca = (a[..., 1] + a[..., 3]) / f32(2)
    ha = a[..., 2] - a[..., 0]
    wa = a[..., 3] - a[..., 1]
    w = np.exp(r[..., 3].astype(np.float64)).astype(f32) * wa
    h = np.exp(r[..., 2].astype(np.float64)).astype(f32) * ha
    yc = r[..., 0] * ha + yca
    xc = r[..., 1] * wa + xca
    ymin = yc - h / f32(2)
    xmin = xc - w / f32(2)
    ymax = yc + h / f32(2)
    xmax = xc + w / f32(2)
    boxes = np.stack([xmin, ymin, xmax, ymax], axis=2).astype(f32)
    boxes[:, :, 0] = np.maximum(boxes[:, :, 0], f32(0))
    boxes[:, :, 1] = np.maximum(boxes[:, :, 1], f32(0))
    boxes[:, :, 2] = np.minimum(boxes[:, :, 2], f32(img_w - 1))
    boxes[:, :, 3] = np.minimum(boxes[:, :, 3], f32(img_h - 1))
    return boxes


def nms(boxes, scores, thr):
    """Greedy NMS, fp32, stable descending order; returns kept indices in that order (int64)."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    b = boxes.astype(f32)
    order = np.argsort(-scores.astype(f32), kind="stable")
    b = b[order]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = f32(thr)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(f32(0), xx2 - xx1)
        h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= ovr > thr
    return order[np.asarray(keep, dtype=np.int64)]


def batched_nms(boxes, scores, idxs, thr, device="cuda", mode=None):
    """torchvision.ops.boxes.batched_nms; ``device`` selects the numel threshold of the dispatch."""
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    trick = boxes.size <= (4000 if device == "cpu" else 100000)
    if mode is not None:
        trick = mode == "trick"
    boxes = boxes.astype(f32)
    if trick:
        max_coordinate = boxes.max()
        offsets = idxs.astype(f32) * (max_coordinate + f32(1))
        return nms(boxes + offsets[:, None], scores, thr)
    keep_mask = np.zeros(scores.shape[0], dtype=bool)
    for c in np.unique(idxs):
        cur = np.where(idxs == c)[0]
        keep_mask[cur[nms(boxes[cur], scores[cur], thr)]] = True
    keep = np.where(keep_mask)[0]
    return keep[np.argsort(-scores[keep].astype(f32), kind="stable")]


def det_postprocess(anchors, regression, classification, img_h, img_w, threshold, iou_threshold, device="cuda", mode=None,
                    pre_boxes=None):
    """postprocess(): list of dicts {rois, class_ids, scores} (empty arrays when nothing survives)."""
    boxes = pre_boxes if pre_boxes is not None else bbox_transform_clip(anchors.reshape(-1, 4), regression, img_h, img_w)
    cls = classification.astype(f32)
    scores = cls.max(axis=2)
    out = []
    for i in range(cls.shape[0]):
        m = scores[i] > f32(threshold)
        if m.sum() == 0:
            out.append({"rois": np.array(()), "class_ids": np.array(()), "scores": np.array(())})
            continue
        cp, bp, sp = cls[i, m], boxes[i, m], scores[i, m]
        classes = cp.argmax(axis=1)
        keep = batched_nms(bp, sp, classes, iou_threshold, device, mode)
        if keep.shape[0] == 0:
            out.append({"rois": np.array(()), "class_ids": np.array(()), "scores": np.array(())})
        else:
            out.append({"rois": bp[keep], "class_ids": classes[keep].astype(np.int64), "scores": sp[keep]})
    return out


def seg_argmax(masks):
    return np.argmax(masks, axis=1).astype(np.int64)


def softmax2(logits):
    x = logits.astype(f32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis=-1, keepdims=True)).astype(f32)


def decode_lane(prob, loc, fh, fw, ppl, step, interval, input_w, input_h, thr, margin=100.0):
    """Candidate lanes of one image in (h, w) scan order.

    prob [fh*fw] fp32 P(lane), loc [fh*fw, 2*ppl+2] fp32.  Returns a list of dicts with the fields of
    the reference's ``Lane`` (prob, start_pos, end_pos, ax, ay, xs fp32 [n], ys float64 [n]).
    """
    lanes = []
    ppa = ppl / fh
    interval = float(interval)
    thr32 = f32(thr)
    for h in range(fh):
        for w in range(fw):
            idx = h * fw + w
            p = f32(prob[idx])
            if p < thr32:
                continue
            y_pos = int((fh - 1 - h) * ppa)
            cx = (1.0 * w + 0.5) * step
            cy = (1.0 * h + 0.5) * step
            up = loc[idx, ppl + 2: 2 * ppl + 2].astype(f32) * f32(interval)
            end_up = f32(loc[idx, ppl + 1])
            ux, uy = [], []
            end_pos = start_pos = y_pos
            for i in range(ppl):
                if f32(i) >= end_up or y_pos + i >= ppl:
                    break
                x = f32(cx) + up[i]
                if x < f32(0) or x >= f32(input_w):
                    break
                ux.append(x)
                uy.append(input_h - 1 - (y_pos + i) * interval)
                end_pos = y_pos + i + 1
            down = loc[idx, :ppl].astype(f32) * f32(interval)
            end_down = f32(loc[idx, ppl])
            dx, dy = [], []
            for i in range(y_pos):
                if f32(i) >= end_down or y_pos - 1 - i < 0:
                    break
                x = f32(cx) + down[i]
                if x < f32(0) or x >= f32(input_w + margin):
                    break
                dx.insert(0, x)
                dy.insert(0, input_h - 1 - (y_pos - 1 - i) * interval)
                start_pos = y_pos - 1 - i
            if len(ux) + len(dx) >= 2:
                lanes.append(dict(prob=p, start_pos=start_pos, end_pos=end_pos, ax=cx, ay=cy, anchor=idx,
                                  xs=np.asarray(dx + ux, dtype=f32), ys=np.asarray(dy + uy, dtype=np.float64)))
    return lanes


def lane_dist(l1, l2, use_mean=False):
    ms = max(l1["start_pos"], l2["start_pos"])
    me = min(l1["end_pos"], l2["end_pos"])
    if me <= ms or ms < 0 or me < 1:
        return 10e6
    x1, x2 = l1["xs"], l2["xs"]
    dis = f32(0)
    for i in range(ms, me):
        dis = f32(dis + abs(f32(x1[i - l1["start_pos"]] - x2[i - l2["start_pos"]])))
    dis = f32(dis / f32(me - ms))
    if use_mean:
        return dis
    dis = max(dis, abs(f32(x1[ms - l1["start_pos"]] - x2[ms - l2["start_pos"]])))
    dis = max(dis, abs(f32(x1[me - 1 - l1["start_pos"]] - x2[me - 1 - l2["start_pos"]])))
    return dis


def lane_nms(lanes, thresh, use_mean=False):
    if len(lanes) == 0:
        return []
    order = sorted(range(len(lanes)), key=lambda i: -float(lanes[i]["prob"]))  # stable, like sorted(lane_set)
    ls = [lanes[i] for i in order]
    sel = [False] * len(ls)
    out = []
    for n in range(len(ls)):
        if sel[n]:
            continue
        out.append(ls[n])
        sel[n] = True
        for t in range(n + 1, len(ls)):
            d = lane_dist(ls[n], ls[t], use_mean)
            if (d <= thresh) if isinstance(d, float) else (d <= f32(thresh)):
                sel[t] = True
    return out


def lane_decode_nms(cls, loc, fh, fw, ppl, step, interval, input_w, input_h, conf_thres, nms_thres, use_mean=False,
                    cls_is_prob=False):
    prob = cls[:, 1].astype(f32) if cls_is_prob else softmax2(cls)[:, 1]
    return lane_nms(decode_lane(prob, loc, fh, fw, ppl, step, interval, input_w, input_h, conf_thres), nms_thres, use_mean)
