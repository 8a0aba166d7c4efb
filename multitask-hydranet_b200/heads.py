"""The three task heads: parameter holders plus the reference's *static* decode API, backed by the
native post-processing kernels (``csrc/hn_postproc.cu``).

Interfaces mirrored (same names, argument meaning, return types and error behaviour):
  SegmentHeader.decode      head_seg/segmentation.py:107-125   (argmax part on the GPU; colouring is host cv2)
  DetectionHeader.decode    head_detect/detection.py:232-245 -> detection_loss.py:70-108
  DetectionHeader.invert_affine  detection.py:217-230
  LaneHeader.decode         head_lane/lanedetect.py:103-116 -> lane_codec.py:116-219, lane_codec_utils.py:518-542
  LaneHeader.scale_to_org   lanedetect.py:118-124
"""
import itertools

import numpy as np
import torch
from torch import nn

from . import _native as nv
from .lane_codec import Lane, LaneCodec, Point, convert_lane_to_dict, order_lane_x_axis
from .modules import _Conv3x3, _ConvBlock, _Tower


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("%s: expected a CUDA tensor -- the B200 path has no CPU fallback" % what)


# ------------------------------------------------------------------------------------------------
class SegmentHeader(nn.Module):
    def __init__(self, num_ch_enc, num_ch_dec=None, num_output_channels=10, use_skips=True):
        super().__init__()
        if not use_skips:
            raise NotImplementedError("use_skips=False is never used by the reference configs")
        self.num_output_channels, self.use_skips = num_output_channels, use_skips
        self.num_ch_enc, self.num_ch_dec = list(num_ch_enc), list(num_ch_dec)
        mods = []
        for i in range(len(num_ch_enc) - 1, -1, -1):
            cin = num_ch_enc[-1] if i == len(num_ch_enc) - 1 else num_ch_dec[i + 1]
            mods.append(_ConvBlock(cin, num_ch_dec[i]))
            cin = num_ch_dec[i] + (num_ch_enc[i - 1] if i > 0 else 0)
            mods.append(_ConvBlock(cin, num_ch_dec[i]))
        mods.append(_Conv3x3(num_ch_dec[0], num_output_channels))
        self.decoder = nn.ModuleList(mods)

    @staticmethod
    def argmax(masks):
        """``torch.argmax(masks, dim=1)`` (int64 [B,H,W]) computed by the native kernel."""
        _require_cuda(masks, "SegmentHeader.argmax")
        m = masks.detach()
        if m.dtype != torch.float32 or not m.is_contiguous():
            m = m.float().contiguous()
        B, Cc, H, W = m.shape
        out = torch.empty((B, H, W), dtype=torch.int64, device=m.device)
        with torch.cuda.device(m.device):
            nv.check(nv.lib.hn_seg_argmax(m.data_ptr(), B, Cc, H * W, out.data_ptr(), None, _stream_ptr(m.device)))
        return out

    @staticmethod
    def decode(imgs, masks, org_size, vis_color_id):
        import cv2
        seg_predictions = SegmentHeader.argmax(masks).cpu().numpy()
        visual_imgs = []
        for batch_idx in range(len(imgs)):
            seg_prediction = seg_predictions[batch_idx]
            vis_seg = np.zeros([seg_prediction.shape[0], seg_prediction.shape[1], 3], dtype=np.uint8)
            for cls_id, color in vis_color_id.items():
                vis_seg[seg_prediction == cls_id] = color
            vis_seg = cv2.resize(vis_seg, org_size, cv2.INTER_NEAREST)
            visual_imgs.append(cv2.addWeighted(imgs[batch_idx], 0.8, vis_seg, 0.5, 0.0))
        return visual_imgs


# ------------------------------------------------------------------------------------------------
def make_anchors(image_shape, anchor_scale, strides, scales, ratios):
    """[1, A, 4] fp32 (y1, x1, y2, x2): level-major, then y, x, then (scale-major, ratio-minor).

    Same float64 arithmetic and ordering as ``Anchors.forward`` (head_detect/detection.py:108-170).
    """
    H, W = int(image_shape[0]), int(image_shape[1])
    levels = []
    for stride in strides:
        if W % stride != 0 or H % stride != 0:
            raise ValueError('input size must be divided by the stride.')
        x = np.arange(stride / 2, W, stride)
        y = np.arange(stride / 2, H, stride)
        xv, yv = np.meshgrid(x, y)
        xv, yv = xv.reshape(-1), yv.reshape(-1)
        per = []
        for scale, ratio in itertools.product(scales, ratios):
            base = anchor_scale * stride * scale
            ax2, ay2 = base * ratio[0] / 2.0, base * ratio[1] / 2.0
            per.append(np.stack([yv - ay2, xv - ax2, yv + ay2, xv + ax2], axis=1))
        levels.append(np.stack(per, axis=1).reshape(-1, 4))
    return np.concatenate(levels, 0).astype(np.float32)[None]


class _Anchors(nn.Module):
    def __init__(self, anchor_scale, pyramid_levels, scales, ratio):
        super().__init__()
        self.anchor_scale = anchor_scale
        self.pyramid_levels = [3, 4, 5, 6, 7] if pyramid_levels is None else pyramid_levels
        self.strides = [2 ** x for x in self.pyramid_levels]
        self.scales, self.ratios = scales, ratio
        self.last_anchors, self.last_shape = {}, None

    def forward(self, image, dtype=torch.float32):
        shape = tuple(image.shape[2:])
        if shape == self.last_shape and image.device in self.last_anchors:
            return self.last_anchors[image.device]
        if self.last_shape != shape:
            self.last_shape, self.last_anchors = shape, {}
        a = make_anchors(shape, self.anchor_scale, self.strides, self.scales, self.ratios)
        if dtype == torch.float16:
            a = a.astype(np.float16)
        t = torch.from_numpy(a).to(image.device)
        self.last_anchors[image.device] = t
        return t


class DetectionHeader(nn.Module):
    def __init__(self, num_classes, fpn_num_filters_detect, aspect_ratios, scales, box_class_repeats, pyramid_levels,
                 anchor_scale, onnx_export=False):
        super().__init__()
        self.num_classes, self.fpn_num_filters_detect = num_classes, fpn_num_filters_detect
        self.aspect_ratios, self.scales = aspect_ratios, scales
        self.num_anchors = len(aspect_ratios) * len(scales)
        self.box_class_repeats, self.pyramid_levels, self.anchor_scale = box_class_repeats, pyramid_levels, anchor_scale
        self.regressor = _Tower(fpn_num_filters_detect, self.num_anchors * 4, box_class_repeats, pyramid_levels)
        self.classifier = _Tower(fpn_num_filters_detect, self.num_anchors * num_classes, box_class_repeats, pyramid_levels)
        self.anchors = _Anchors(anchor_scale, (torch.arange(pyramid_levels) + 3).tolist(), scales, aspect_ratios)

    @staticmethod
    def invert_affine(metas, preds):
        for i in range(len(preds)):
            if len(preds[i]['rois']) == 0:
                continue
            if metas is float:
                preds[i]['rois'][:, [0, 2]] = preds[i]['rois'][:, [0, 2]] / metas
                preds[i]['rois'][:, [1, 3]] = preds[i]['rois'][:, [1, 3]] / metas
            else:
                new_w, new_h, old_w, old_h, padding_w, padding_h = metas[i]
                preds[i]['rois'][:, [0, 2]] = preds[i]['rois'][:, [0, 2]] / (new_w / old_w)
                preds[i]['rois'][:, [1, 3]] = preds[i]['rois'][:, [1, 3]] / (new_h / old_h)
        return preds

    @staticmethod
    def decode_device(img_hw, regressions, classifications, anchors, conf_thres=0.6, iou_thres=0.3,
                      nms_mode=nv.NMS_AUTO_CUDA, pre_boxes=None, workspace=None):
        """Device-side result: (boxes [N,A,4], scores [N,A], class_ids [N,A] int64, count [N] int32, cand [N]).

        Rows ``[:count[n]]`` of image n are the kept detections in score-descending order.
        """
        _require_cuda(classifications, "DetectionHeader.decode")
        cls = classifications.detach().float().contiguous()
        N, A, ncls = cls.shape
        dev = cls.device
        reg = regressions.detach().float().contiguous() if regressions is not None else None
        anc = anchors.detach().float().contiguous().view(-1, 4) if anchors is not None else None
        if anc is not None and anc.shape[0] != A:
            raise ValueError("anchors (%d) do not match the predictions (%d)" % (anc.shape[0], A))
        pre = pre_boxes.detach().float().contiguous() if pre_boxes is not None else None
        nbytes = nv.lib.hn_det_workspace_bytes(N, A)
        if workspace is None or workspace.numel() < nbytes:
            workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)  # zeroed: the rounds kernel's masked int4 loads read past list ends
        boxes = torch.empty((N, A, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((N, A), dtype=torch.float32, device=dev)
        cids = torch.empty((N, A), dtype=torch.int64, device=dev)
        count = torch.zeros((max(N, 1),), dtype=torch.int32, device=dev)
        cand = torch.zeros((max(N, 1),), dtype=torch.int32, device=dev)
        d = nv.DetDesc(anc.data_ptr() if anc is not None else None, reg.data_ptr() if reg is not None else None,
                       cls.data_ptr(), N, A, ncls, int(img_hw[0]), int(img_hw[1]), float(conf_thres), float(iou_thres),
                       int(nms_mode), workspace.data_ptr(), workspace.numel(), boxes.data_ptr(), scores.data_ptr(),
                       cids.data_ptr(), count.data_ptr(), cand.data_ptr(), pre.data_ptr() if pre is not None else None)
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_det_decode_nms(d, _stream_ptr(dev)))
        return boxes, scores, cids, count[:N], cand[:N]

    @staticmethod
    def invert_affine_device(metas, boxes, count):
        """``invert_affine`` (detection.py:218-230) on the device, in place on ``decode_device``'s boxes [N, A, 4] (rows
        ``[:count[n]]``): same fp32 divisions as numpy performs on the host (float32 rois / Python float).  ``metas``: a float, or
        per image (new_w, new_h, old_w, old_h, padding_w, padding_h)."""
        _require_cuda(boxes, "DetectionHeader.invert_affine")
        N, A = boxes.shape[0], boxes.shape[1]
        if isinstance(metas, float):
            sc = np.full((N, 2), np.float32(metas), dtype=np.float32)
        else:
            sc = np.array([[np.float32(m[0] / m[2]), np.float32(m[1] / m[3])] for m in metas], dtype=np.float32)
        scale = torch.from_numpy(sc).to(boxes.device)
        with torch.cuda.device(boxes.device):
            nv.check(nv.lib.hn_det_invert_affine(boxes.data_ptr(), count.data_ptr(), N, A, scale.data_ptr(), _stream_ptr(boxes.device)))
        return boxes

    @staticmethod
    def decode(imgs, regressions, classifications, anchors, conf_thres=0.6, iou_thres=0.3, metas=None):
        """detection.py:232-245.  ``metas`` (extension): apply ``invert_affine`` on the device before the download, so the
        returned rois are in original-image coordinates."""
        if imgs is None:
            return None
        boxes, scores, cids, count, _ = DetectionHeader.decode_device(imgs.shape[2:], regressions, classifications, anchors,
                                                                      conf_thres, iou_thres)
        if metas is not None:
            DetectionHeader.invert_affine_device(metas, boxes, count)
        counts = count.cpu().tolist()  # the one host sync the reference also has (.cpu())
        kmax = max(counts) if counts else 0
        b, s, c = boxes[:, :kmax].cpu().numpy(), scores[:, :kmax].cpu().numpy(), cids[:, :kmax].cpu().numpy()
        out = []
        for i, k in enumerate(counts):
            if k == 0:
                out.append({'rois': np.array(()), 'class_ids': np.array(()), 'scores': np.array(())})
            else:
                out.append({'rois': b[i, :k].copy(), 'class_ids': c[i, :k].copy(), 'scores': s[i, :k].copy()})
        return out

    @staticmethod
    def class_color(index, n):
        """BGR colour of class ``index`` of ``n``: evenly spaced hues (the reference looks CSS colour names up through
        ``webcolors``, display.py:35-51 -- cosmetic, and that package is not a dependency here)."""
        import colorsys
        r, g, b = colorsys.hsv_to_rgb((index % max(n, 1)) / float(max(n, 1)), 0.75, 1.0)
        return int(b * 255), int(g * 255), int(r * 255)

    @staticmethod
    def display(decode, imgs, obj_list, org_size, target_size):
        """Draw the decoded boxes on the original-size frames (detection.py:247-252 -> display.py:53-84).

        ``decode``: list of {'rois','class_ids','scores'} in network-input pixels; boxes are truncated to int, scaled by
        org_size / target_size and drawn with a '<class><score %>' tag.  As in the reference, the list is returned after
        the first frame that has detections (display.py:84 returns inside its loop; demo.py runs batch 1), and frames
        without detections are returned untouched."""
        import cv2
        sx, sy = org_size[0] / float(target_size[0]), org_size[1] / float(target_size[1])
        for i in range(len(imgs)):
            rois = decode[i]['rois']
            if len(rois) == 0:
                continue
            canvas = imgs[i] = imgs[i].copy()
            thick = int(round(0.003 * max(canvas.shape[0:2]))) or 1
            font_thick, font_scale = max(thick - 2, 1), float(thick) / 3
            for box, cid, score in zip(rois, decode[i]['class_ids'], decode[i]['scores']):
                x1, y1, x2, y2 = (int(v) for v in box)
                p1 = (int(x1 * sx), int(y1 * sy))
                p2 = (int(x2 * sx), int(y2 * sy))
                name = obj_list[int(cid)]
                colour = DetectionHeader.class_color(obj_list.index(name), len(obj_list))
                cv2.rectangle(canvas, p1, p2, colour, thickness=thick)
                pct = '{:.0%}'.format(float(score))
                (tw, th), _ = cv2.getTextSize(name, 0, fontScale=font_scale, thickness=font_thick)
                (sw, _), _ = cv2.getTextSize(pct, 0, fontScale=font_scale, thickness=font_thick)
                cv2.rectangle(canvas, p1, (p1[0] + tw + sw + 15, p1[1] - th - 3), colour, -1)
                cv2.putText(canvas, name + pct, (p1[0], p1[1] - 2), 0, font_scale, [0, 0, 0], thickness=font_thick,
                            lineType=cv2.FONT_HERSHEY_SIMPLEX)
            return imgs
        return imgs


# ------------------------------------------------------------------------------------------------
class LaneHeader(nn.Module):
    def __init__(self, base_channel, num_classes, stride, input_width, input_height, interval):
        super().__init__()
        self.base_channel, self.num_classes, self.stride = base_channel, num_classes, stride
        self.input_width, self.input_height, self.interval = input_width, input_height, interval
        self.feat_width, self.feat_height = int(input_width / stride), int(input_height / stride)
        self.points_per_line = int(input_height / interval)
        self.lane_up_pts_num = self.points_per_line + 1
        self.lane_down_pts_num = self.points_per_line + 1

        def branch(cout):
            return nn.Sequential(nn.Conv2d(base_channel, base_channel, 1, bias=False), nn.BatchNorm2d(base_channel),
                                 nn.ReLU(inplace=True), nn.Conv2d(base_channel, cout, 1))

        self.conv_cls_conv = branch(num_classes)
        self.conv_up_conv = branch(self.lane_up_pts_num)
        self.conv_down_conv = branch(self.lane_down_pts_num)

    @property
    def input_shape(self):
        return self.feat_height, self.feat_width

    @staticmethod
    def decode_device(predict_cls, predict_loc, pointlane, conf_thres=0.5, nms_line_thres=100, use_mean=False,
                      cls_is_prob=False):
        """Batched device-side decode: tensors [N, na, ...]; returns (count, meta, prob, x) on the device."""
        _require_cuda(predict_cls, "LaneHeader.decode")
        cls = predict_cls.detach().float().contiguous()
        loc = predict_loc.detach().float().contiguous()
        if cls.dim() == 2:
            cls, loc = cls[None], loc[None]
        N, na = cls.shape[0], cls.shape[1]
        fh, fw, ppl = pointlane.feature_height, pointlane.feature_width, pointlane.points_per_line
        if na != fh * fw or loc.shape[2] != 2 * ppl + 2:
            raise ValueError("lane predictions do not match the codec geometry")
        dev = cls.device
        ws = torch.empty(nv.lib.hn_lane_workspace_bytes(N, na, ppl), dtype=torch.uint8, device=dev)
        count = torch.zeros((N,), dtype=torch.int32, device=dev)
        cand = torch.zeros((N,), dtype=torch.int32, device=dev)
        meta = torch.zeros((N, na, 4), dtype=torch.int32, device=dev)
        prob = torch.zeros((N, na), dtype=torch.float32, device=dev)
        xs = torch.zeros((N, na, ppl), dtype=torch.float32, device=dev)
        d = nv.LaneDesc(cls.data_ptr(), loc.data_ptr(), N, fh, fw, ppl, int(bool(cls_is_prob)),
                        float(np.float32(conf_thres)), float(np.float32(nms_line_thres)), int(bool(use_mean)),
                        float(pointlane.step_w), float(pointlane.interval), float(pointlane.points_per_anchor),
                        float(pointlane.input_width), 100.0, ws.data_ptr(), count.data_ptr(), meta.data_ptr(),
                        prob.data_ptr(), xs.data_ptr(), cand.data_ptr())
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_lane_decode_nms(d, _stream_ptr(dev)))
        return count, meta, prob, xs, cand

    @staticmethod
    def lanes_from_device(count, meta, prob, xs, pointlane, index=0):
        """Rebuild the reference's ``Lane`` objects (numpy float32 x, Python float y) for one image."""
        k = int(count[index])
        meta, prob, xs = meta[index, :k].cpu().numpy(), prob[index, :k].cpu().numpy(), xs[index, :k].cpu().numpy()
        lanes = []
        for i in range(k):
            a, start, end, npts = (int(v) for v in meta[i])
            h, w = divmod(a, pointlane.feature_width)
            pts = np.empty(npts, dtype=object)
            for q in range(npts):
                y = pointlane.input_height - 1 - (start + q) * pointlane.interval
                pts[q] = Point(xs[i, q], y)
            lanes.append(Lane(prob[i], start, end, (1.0 * w + 0.5) * pointlane.step_w, (1.0 * h + 0.5) * pointlane.step_h, 1, pts))
        return lanes

    @staticmethod
    def decode(predict_cls, predict_loc, pointlane, conf_thres=0.5, nms_line_thres=100, use_mean=False):
        count, meta, prob, xs, _ = LaneHeader.decode_device(predict_cls, predict_loc, pointlane, conf_thres, nms_line_thres, use_mean)
        return LaneHeader.lanes_from_device(count.cpu(), meta, prob, xs, pointlane, 0)

    @staticmethod
    def scale_to_org(lane_nms_set, net_input_width, net_input_height, org_width, org_height):
        lane_order_set = order_lane_x_axis(list(lane_nms_set), net_input_height)
        return convert_lane_to_dict(lane_order_set, org_width / net_input_width, org_height / net_input_height)

    @staticmethod
    def scale_to_org_device(count, meta, prob, xs, pointlane, net_input_width, net_input_height, org_width, org_height, index=0):
        """``scale_to_org(decode(...))`` without rebuilding ``Lane`` objects on the host: the ordering keys (slope, crossing of the
        bottom row) and the points in original-frame coordinates are computed by ``hn_lane_scale_to_org`` from ``decode_device``'s
        tensors; the host only runs the reference's (non-transitive) comparator through Python's sort and builds the dict.
        Returns exactly what ``scale_to_org`` returns (lanedetect.py:118-124)."""
        _require_cuda(xs, "LaneHeader.scale_to_org")
        N, na, ppl = xs.shape
        dev = xs.device
        keys = torch.empty((N, na, 4), dtype=torch.float32, device=dev)
        xo = torch.empty((N, na, ppl), dtype=torch.float32, device=dev)
        yo = torch.empty((N, na, ppl), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            nv.check(nv.lib.hn_lane_scale_to_org(count.data_ptr(), meta.data_ptr(), xs.data_ptr(), N, na, ppl, float(pointlane.input_height),
                                                 float(pointlane.interval), float(net_input_height - 1.0), org_width / net_input_width,
                                                 org_height / net_input_height, keys.data_ptr(), xo.data_ptr(), yo.data_ptr(), _stream_ptr(dev)))
        k = int(count[index])
        meta_h, prob_h = meta[index, :k].cpu().numpy(), prob[index, :k].cpu().numpy()
        keys_h, xo_h, yo_h = keys[index, :k].cpu().numpy(), xo[index, :k].cpu().numpy(), yo[index, :k].cpu().numpy()

        class _Key:  # LaneWithCrossK.__lt__ (lane_codec_utils.py:168-182) on the device-computed keys
            __slots__ = ("i", "cross_x", "k", "last_x")

            def __init__(self, i):
                self.i, self.cross_x, self.k, self.last_x = i, keys_h[i, 0], keys_h[i, 1], keys_h[i, 3]

            def __lt__(self, other):
                if abs(self.cross_x - other.cross_x) > 2.0:
                    return self.cross_x < other.cross_x
                return self.last_x < other.last_x

        order = sorted(_Key(i) for i in range(k))
        lines = []
        for key in order:
            i = key.i
            if prob_h[i] < 0.01:
                continue
            n = int(meta_h[i, 3])
            lines.append({'score': prob_h[i], 'points': [{'x': xo_h[i, q], 'y': float(yo_h[i, q])} for q in range(n)]})
        return {'Lines': lines}

    @staticmethod
    def visual(imgs, predict_jsons, org_width=1920, min_length=2, filter_vertical=True, filter_thres=65):
        """Draw lanes (``scale_to_org(...)["Lines"]`` per frame) in place on the frames (lanedetect.py:126-178): lines shorter
        than ``min_length`` points are skipped, so are -- with ``filter_vertical`` -- lines whose least-squares slope is
        steeper than ``filter_thres`` degrees; each lane is a 15-px polyline plus a 'Lane: <score>' tag."""
        import cv2
        out = []
        for frame, lines in zip(imgs, predict_jsons):
            for line in lines:
                pts = [(int(p["x"]), int(p["y"])) for p in line["points"]]
                if len(pts) < min_length:
                    continue
                if filter_vertical:
                    arr = np.array(pts)
                    slope = np.polyfit(arr[:, 0], arr[:, 1], 1)[0]
                    if abs(np.arctan(slope)) / 3.1415 * 180 > filter_thres:
                        continue
                for a, b in zip(pts[:-1], pts[1:]):
                    frame = cv2.line(frame, a, b, color=(255, 255, 0), thickness=15)
                tx, ty = pts[min_length - 1]
                if tx < 0:
                    tx = 30
                if tx > org_width:
                    tx, ty = org_width - 300, ty - 60
                cv2.putText(frame, "%s: %.2f" % ("Lane", float(line["score"])), (tx, ty - 10), cv2.FONT_HERSHEY_SIMPLEX, 2.0,
                            (255, 255, 0), 7)
            out.append(frame)
        return out
