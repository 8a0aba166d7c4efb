"""Drop-in ``HydraNet`` facade (reference: model/model.py:26-264).

Same constructor, attribute names, ``state_dict`` keys, ``forward(x, mode)`` signature and output
structure as the reference.  In eval mode the forward runs entirely in the native sm_100a engine
(``engine.py`` -> ``libhydranet_b200.so``); in train mode it runs as autograd functions over the same
native kernels (``train.py``: batch-statistics BatchNorm, dgrad / wgrad on tcgen05).  There is no
PyTorch/cuDNN or CPU fallback: a missing library fails at import, a CPU tensor raises.

Output lifetime: by default ``forward`` returns fresh tensors (as the reference does).  With
``model.static_outputs = True`` (zero-copy serving) the returned tensors are the plan's static buffers and
are overwritten by the next forward of the same shape.
"""
import collections
import os

import torch
from torch import nn

from . import _native as nv
from .engine import Plan, SplitPlan
from .heads import DetectionHeader, LaneHeader, SegmentHeader
from .modules import RegNetY, StackBiFPN


class HydraNet(nn.Module):
    def __init__(self, cfgs, onnx_export=False):
        super().__init__()
        self.cfgs, self.onnx_export = cfgs, onnx_export
        self.net_input_width = cfgs["dataloader"]["network_input_width"]
        self.net_input_height = cfgs["dataloader"]["network_input_height"]
        bb = cfgs["backbone"]
        self.backbone = RegNetY(bb["initial_width"], bb["slope"], bb["quantized_param"], bb["network_depth"],
                                bb["bottleneck_ratio"], bb["group_width"], bb["stride"], bb["se_ratio"])
        self.fpn_num_filters, self.fpn_cell_repeats = bb["fpn_num_filters"], bb["fpn_cell_repeats"]
        self.conv_channel_coef = bb["conv_channel_coef"]
        self.neck = StackBiFPN(self.fpn_num_filters, self.fpn_cell_repeats, self.conv_channel_coef)

        self.train_detect = cfgs["train"]["train_detect"]
        if self.train_detect:
            dc = cfgs["detection"]
            self.num_classes = dc["num_classes"]
            r1, r2 = dc["aspect_ratios_factor"]
            self.aspect_ratios = [(1.0, 1.0), (r1, r2), (r2, r1)]
            self.scales = [2 ** s for s in dc["scales_factor"]]
            self.detectheader = DetectionHeader(dc["num_classes"], dc["fpn_num_filters_detect"], self.aspect_ratios,
                                                self.scales, dc["box_class_repeats"], dc["pyramid_levels"],
                                                dc["anchor_scale"], onnx_export)
        else:
            self.detectheader = None

        self.train_seg = cfgs["train"]["train_seg"]
        if self.train_seg:
            sc = cfgs["segment"]
            self.segment_class_list = sc["class_list"]
            self.segheader = SegmentHeader(sc["channel_dimension_seg_encode"], sc["channel_dimension_seg_decode"],
                                           len(sc["class_list"]))
        else:
            self.segheader = None

        self.train_lane = cfgs["train"]["train_lane"]
        if self.train_lane:
            lc = cfgs["lane"]
            self.laneheader = LaneHeader(lc["base_channel"], lc["num_classes"], lc["anchor_stride"], self.net_input_width,
                                         self.net_input_height, lc["interval"])
        else:
            self.laneheader = None
        self.loss_detect = self.loss_seg = self.loss_cls = self.loss_reg = None
        self._plans = collections.OrderedDict()
        self.max_plans = 4          # LRU bound: every plan owns a full activation set, packed weights and a CUDA graph
        self._dirty = True          # parameters may have changed since the plans were packed
        self._sig = None
        self._fused_post = None
        self._last_plan = None
        object.__setattr__(self, "_shadow", None)
        self._graph_stream = {}
        self.use_graph = False
        # False (default): forward returns fresh tensors, like the reference.  True: zero-copy serving mode -- the returned
        # tensors are the plan's static buffers, valid until the next forward of the same shape.
        self.static_outputs = False
        # run the three heads as independent branches of the plan (forked streams / a forked CUDA graph)
        self.head_branches = os.environ.get("HN_BRANCHES", "1") != "0"
        # batches >= 4 as two half-batch plans interleaved on two streams (engine.SplitPlan).  Off by default: measured
        # 9.67 vs 9.40 ms/step at batch 32 -- the persistent conv CTAs take a whole SM's shared memory, so kernels of the
        # two halves cannot share SMs and every launch's fixed cost is simply paid twice.
        self.split_batch = os.environ.get("HN_SPLIT", "0") == "1"

    # -- native engine management ------------------------------------------------------------
    # Packed weights are rebuilt when the parameters change.  Walking all 1 177 tensors costs ~1 ms of Python, so it is
    # done only after an event that can change them: load_state_dict, .to()/.cuda()/.half() (_apply), a train()/eval()
    # switch (an optimizer may have stepped in between), or an explicit refresh().
    def refresh(self):
        """Tell the engine that parameters / buffers were modified in place (e.g. ``p.data.mul_()`` in eval mode)."""
        self._dirty = True

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._dirty = True
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    def train(self, mode=True):
        self._dirty = True
        return super().train(mode)

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def plan(self, B, H, W, device):
        if self._dirty:
            sig = self._signature()
            if sig != self._sig:  # weights changed: drop every packed plan
                self._plans.clear()
                self._sig = sig
                self._shadow = None
            self._dirty = False
        key = (B, H, W, str(device), self._fused_key(), bool(self.head_branches), bool(self.split_batch))
        plan = self._plans.get(key)
        if plan is None:
            with torch.no_grad(), torch.cuda.device(device):
                split = self.split_batch and B >= 4 and self._fused_post is None
                plan = (SplitPlan if split else Plan)(self, B, H, W, device)
            self._plans[key] = plan
            while len(self._plans) > max(1, int(self.max_plans)):
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return plan

    def packing_source(self):
        """The module tree the schedule builder reads weights from: a CPU copy of this model when the parameters live on
        the GPU, so BatchNorm folding / weight packing is host arithmetic plus one upload per packed matrix (no swarm of
        tiny torch kernels on the device; plan building launches nothing but buffer fills)."""
        if all(not p.is_cuda for p in self.parameters()):
            return self
        if getattr(self, "_shadow", None) is None:
            shadow = HydraNet(self.cfgs, self.onnx_export)
            shadow.load_state_dict({k: v.detach().cpu() for k, v in self.state_dict().items()})
            shadow.eval()
            object.__setattr__(self, "_shadow", shadow)  # not a registered sub-module: invisible to state_dict / .to()
        self._shadow._fused_post = self._fused_post
        self._shadow.head_branches = self.head_branches
        return self._shadow

    def input_buffer(self, B, H, W, device):
        """The plan's static fp32 [B,3,H,W] input.  A caller that writes its batch there (e.g. ``preprocess(..., out=buf)``)
        and passes the same tensor to ``forward`` skips the device-to-device input copy."""
        return self.plan(B, H, W, torch.device(device)).x

    def forward(self, x, mode="train"):
        if not x.is_cuda:
            raise RuntimeError("HydraNet.forward: input must be a CUDA tensor -- the B200 path has no CPU fallback")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input of shape [B, 3, H, W]")
        B, _, H, W = x.shape
        if self.train_detect:
            for s in self.detectheader.anchors.strides:
                if W % s != 0 or H % s != 0:
                    raise ValueError('input size must be divided by the stride.')
        if self.training:
            from .train import train_forward
            with torch.cuda.device(x.device):
                return train_forward(self, x, mode)
        with torch.cuda.device(x.device):  # native launches, side streams and events all use the CURRENT device
            return self._eval_forward(x, mode, B, H, W)

    def _eval_forward(self, x, mode, B, H, W):
        dev = x.device
        plan = self.plan(B, H, W, dev)
        cur = torch.cuda.current_stream(dev)
        run_on = cur
        if self.use_graph and cur.cuda_stream == 0:
            # stream capture is illegal on the legacy default stream: replay on a model-owned stream, forked from and
            # joined back into the caller's stream
            run_on = self._graph_stream.get(dev)
            if run_on is None:
                run_on = self._graph_stream[dev] = torch.cuda.Stream(dev)
        if x.data_ptr() != plan.x.data_ptr():
            plan.x.copy_(x)
        if run_on is not cur:
            run_on.wait_stream(cur)
        stream = run_on.cuda_stream
        with torch.cuda.stream(run_on):
            if self.use_graph:
                if not plan.graph_ready:
                    plan.run(stream)  # warm-up outside capture (lazy function attributes)
                    plan.capture(stream)
                plan.launch_graph(stream)
            else:
                plan.run(stream)
        if run_on is not cur:
            cur.wait_stream(run_on)
        stream = cur.cuda_stream
        o = plan.out
        self._last_plan = plan
        fresh = (lambda t: t) if self.static_outputs else (lambda t: t.clone())
        output_dict = {}
        if self.train_seg:
            output_dict["seg"] = fresh(o["seg"])
        anchors = regression = classification = lane_cls = lane_reg = None
        if self.train_detect:
            anchors = self.detectheader.anchors(x, x.dtype)
            regression, classification = fresh(o["regression"]), fresh(o["classification"])
            output_dict["detection"] = {"anchors": anchors, "regression": regression, "classification": classification}
        if self.train_lane:
            lane_cls, lane_reg = fresh(o["predict_cls"]), fresh(o["predict_loc"])
            output_dict["lane"] = dict(predict_cls=lane_cls, predict_loc=lane_reg)
        if mode != "deploy":
            return output_dict
        seg_cls = None
        if self.train_seg:  # fused arg-max of the last seg conv (model.py:197)
            u8 = o["seg_cls_u8"]
            seg_cls = torch.empty(u8.shape, dtype=torch.int64, device=u8.device)
            nv.check(nv.lib.hn_u8_to_i64(u8.data_ptr(), seg_cls.data_ptr(), u8.numel(), stream))
        return seg_cls, anchors, regression, classification, lane_cls, lane_reg

    def fuse_postprocess(self, det=None, lane=None):
        """Serving mode: run the decoders INSIDE forward, in the detection / lane branches of the plan, so that they overlap
        the segmentation head.  det = (conf_thres, iou_thres) as DetectionHeader.decode takes them; lane = (LaneCodec,
        conf_thres, nms_line_thres, use_mean) as LaneHeader.decode.  ``forward`` returns what it always returns;
        ``postprocess_results()`` hands out the decoders' device tensors (same tuples as ``decode_device``).
        Call with no arguments to switch it off."""
        self._fused_post = {"det": tuple(det) if det else None, "lane": tuple(lane) if lane else None} if (det or lane) else None

    def _fused_key(self):
        f = self._fused_post
        if not f:
            return None
        lane = f.get("lane")
        if lane:
            c = lane[0]
            lane = (c.feature_height, c.feature_width, c.points_per_line, float(c.step_w), float(c.interval)) + tuple(lane[1:])
        return (f.get("det"), lane)

    def postprocess_results(self):
        """(detections, lanes) of the last forward: the tuples DetectionHeader.decode_device / LaneHeader.decode_device return.
        They are the plan's static buffers: valid until the next forward of the same shape (clone to keep them)."""
        o = self._last_plan.out
        return (o["det_post"].result() if "det_post" in o else None), (o["lane_post"].result() if "lane_post" in o else None)

    def seg_class_map(self):
        """uint8 [B,H,W] arg-max of the last forward's seg logits (fused into the final conv's epilogue)."""
        return self._last_plan.out["seg_cls_u8"]

    def cal_loss(self, pred_dict, gt_dict):
        """model/model.py:201-264: the three heads' losses on the forward's outputs (PyTorch autograd on the fp32 head
        tensors; the loss arithmetic lives in ``losses.py``)."""
        from .losses import cal_loss
        return cal_loss(self, pred_dict, gt_dict)
