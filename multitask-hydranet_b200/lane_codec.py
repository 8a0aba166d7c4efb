"""Decode-side lane geometry: the host objects the reference's callers receive.

``LaneCodec`` keeps the constructor and attributes of head_lane/lane_codec.py:10-51 (the GT encoder,
lane_codec.py:53-114, is training-data preparation and out of scope).  ``decode_lane`` runs the
native decode kernel *without* NMS semantics changes: it returns every candidate lane in (h, w)
scan order exactly like lane_codec.py:116-219.  ``Point`` / ``Lane`` / ``order_lane_x_axis`` /
``convert_lane_to_dict`` restate lane_codec_utils.py:5-64, 86-126, 185-282 (the host tail after the
kernel).
"""
import numpy as np


class Point:
    def __init__(self, x=0, y=0):
        self.x, self.y = x, y

    def __repr__(self):
        return "{}, {}".format(self.x, self.y)


class Lane:
    def __init__(self, prob=0, start_pos=0, end_pos=0, anchor_x=0, anchor_y=0, type=0, lane=np.array([])):
        self.prob, self.start_pos, self.end_pos, self.lane = prob, start_pos, end_pos, lane
        self.idx, self.ax, self.ay, self.type = 0, anchor_x, anchor_y, type

    def __lt__(self, other):  # sorted() puts the most probable lane first
        return self.prob > other.prob


class LaneCodec(object):
    def __init__(self, input_width, input_height, anchor_stride, points_per_line, do_interpolate=False,
                 anchor_lane_num=1, scale_invariance=True):
        self.input_width, self.input_height, self.stride = input_width, input_height, anchor_stride
        self.feature_width = int(input_width / anchor_stride)
        self.feature_height = int(input_height / anchor_stride)
        self.points_per_line = points_per_line
        self.pt_nums_single_lane = 2 * points_per_line + 2
        self.points_per_anchor = points_per_line / self.feature_height
        self.interval = float(input_height) / points_per_line
        self.feature_size = self.feature_width * self.feature_height
        self.img_center_x = input_width / 2
        self.step_w = self.step_h = anchor_stride
        self.anchor_lane_num, self.interpolation, self.scale_invariance = anchor_lane_num, do_interpolate, scale_invariance

    def encode_lane(self, lane_object, org_width, org_height):
        raise NotImplementedError("GT encoding is training-data preparation (SURVEY.md section 2.1 row 9)")

    def decode_lane(self, predict_type, predict_loc, exist_threshold=0.5, margin_width=100.0):
        """All candidate lanes of one image, in (h, w) scan order; ``predict_type`` holds probabilities."""
        from .heads import LaneHeader
        if not self.scale_invariance:
            raise NotImplementedError("scale_invariance=False is never configured by the reference")
        if margin_width != 100.0:
            raise NotImplementedError("margin_width is fixed to the reference default (100.0)")
        # nms threshold below every possible distance and no overlap rule -> nothing is suppressed
        count, meta, prob, xs, _ = LaneHeader.decode_device(predict_type, predict_loc, self, exist_threshold, -1.0, False,
                                                            cls_is_prob=True)
        lanes = LaneHeader.lanes_from_device(count.cpu(), meta, prob, xs, self, 0)
        lanes.sort(key=lambda l: (round(l.ay / self.step_h - 0.5), round(l.ax / self.step_w - 0.5)))
        return lanes


def _calc_y_cross(p1, p2, y):
    if abs(p1.y - p2.y) < 1e-6:
        return -1
    k = (p1.x - p2.x) / (p1.y - p2.y)
    b = p1.x - k * p1.y
    return k * y + b


class _LaneWithCrossK:
    def __init__(self, lane_, idx_in, y_in):
        self.lane, self.idx, self.y = lane_, idx_in, y_in
        pts = lane_.lane
        if pts[1].y < pts[0].y:
            self.k = (pts[1].x - pts[0].x) / (pts[1].y - pts[0].y)
            self.cross_x = _calc_y_cross(pts[0], pts[1], y_in)
        elif pts[1].y > pts[0].y:
            self.k = (pts[-1].x - pts[-2].x) / (pts[-1].y - pts[-2].y)
            self.cross_x = _calc_y_cross(pts[-2], pts[-1], y_in)
        else:
            self.k = 1000
            self.cross_x = _calc_y_cross(pts[-2], pts[-1], y_in)

    def __lt__(self, other):
        if abs(self.cross_x - other.cross_x) > 2.0:
            return self.cross_x < other.cross_x
        if self.lane.lane[1].y < self.lane.lane[0].y:
            return self.lane.lane[-1].x < other.lane.lane[-1].x
        return self.lane.lane[0].x < other.lane.lane[0].x


def order_lane_x_axis(lane_set, h):
    """Left-to-right ordering with signed lane indices (lane_codec_utils.py:185-233)."""
    if len(lane_set) == 0:
        return list()
    ordered = sorted(_LaneWithCrossK(l, i, h - 1.0) for i, l in enumerate(lane_set))
    right_pos = len(ordered)
    for i, lw in enumerate(ordered):
        if lw.k > 0:
            right_pos = i
            break
    out = []
    for i, lw in enumerate(ordered):
        lw.lane.idx = (i - right_pos) if i < right_pos else (i - right_pos + 1)
        out.append(lw.lane)
    return out


def convert_lane_to_dict(lane_set, sx, sy):
    lines = []
    for lane in lane_set:
        if lane.prob < 0.01:
            continue
        lines.append({'score': lane.prob, 'points': [{'x': p.x * sx, 'y': p.y * sy} for p in lane.lane]})
    return {'Lines': lines}
