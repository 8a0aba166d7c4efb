"""Schedule builder for the native HydraNet forward.

Turns a ``HydraNet`` parameter tree into (a) packed weights -- BatchNorm folded, K-major bf16
matrices whose K columns follow a per-layer *tap table* -- and (b) a flat list of operator specs
(``ConvSpec`` / ``NodeSpec`` / ...) over NHWC bf16 buffers, which ``Plan.compile`` hands to the C
ABI (``hn_plan_add_*``) once; a forward is then a single ``hn_plan_run``.

Graph being scheduled (reference file:line, /root/reference/model):
  backbone  net/anynet.py:136-145 (Stem 8-20, XBlock 64-76)
  neck      net/bifpn.py:156-233, net/common.py:76-151
  seg head  head_seg/segmentation.py:84-105
  detect    head_detect/detection.py:28-83, 211-215
  lane      head_lane/lanedetect.py:66-96

Design notes:
  * a 3x3 conv over ``cat(up2(X), S)`` never materialises the up-sampled tensor: per output parity
    (py, px) the up-sampled part collapses to a 2x2 conv over X with pre-summed weights, the skip part
    reads S through stride-2 phase views; reflect-pad of up2(X) == replicate-pad of X.
  * grouped 3x3 convs (8 channels / group) run on the tensor cores as block-diagonal 64x64 tiles.
  * zero padding comes for free from TMA out-of-bounds fill; reflect / replicate halos are written by
    the producing conv's epilogue into 1-pixel padded buffers.
"""
import math
import os

import torch

from . import _native as nv

#: squeeze-excite of the small maps (stages 2-4) as ONE cluster launch per block instead of pool+FC -> scale (HN_SE_FUSED=0: off)
SE_FUSED = os.environ.get("HN_SE_FUSED", "1") != "0"
#: ... and, for stride-1 blocks, with the block's grouped 3x3 convolution in front, in the same launch (HN_GCONV_SE=0: off)
GCONV_SE = SE_FUSED and os.environ.get("HN_GCONV_SE", "1") != "0"


# ------------------------------------------------------------------------------------------------
# views and buffers
# ------------------------------------------------------------------------------------------------
class V:
    """4-D NHWC view (element strides) into a tensor's storage."""

    def __init__(self, t, off, N, H, W, C, sn, sy, sx):
        self.t, self.off, self.N, self.H, self.W, self.C, self.sn, self.sy, self.sx = t, off, N, H, W, C, sn, sy, sx

    def to_c(self):
        return nv.View(self.t.data_ptr() + self.t.element_size() * self.off, self.N, self.H, self.W, self.C,
                       self.sn, self.sy, self.sx)

    def chan(self, c0, c):
        return V(self.t, self.off + c0, self.N, self.H, self.W, c, self.sn, self.sy, self.sx)

    def phase(self, ry, rx):
        """rows ry, ry+2, ... and columns rx, rx+2, ... of this view."""
        return V(self.t, self.off + ry * self.sy + rx * self.sx, self.N, (self.H - ry + 1) // 2, (self.W - rx + 1) // 2,
                 self.C, self.sn, 2 * self.sy, 2 * self.sx)

    def flat(self):
        """[1,1,N*H*W,C] row view; only for views whose pixels are equally spaced."""
        assert self.sy == self.W * self.sx and self.sn == self.H * self.sy, "view is not flattenable"
        return V(self.t, self.off, 1, 1, self.N * self.H * self.W, self.C, 0, 0, self.sx)

    def is_flattenable(self):
        return self.sy == self.W * self.sx and self.sn == self.H * self.sy

    def torch_view(self):
        return self.t.view(-1).as_strided((self.N, self.H, self.W, self.C), (self.sn, self.sy, self.sx, 1), self.off)


class Buf:
    """NHWC activation buffer with an optional 1-pixel halo (zero-initialised once)."""

    def __init__(self, device, dtype, N, H, W, C, pad=0, halo=nv.HALO_NONE):
        self.N, self.H, self.W, self.C, self.pad, self.halo = N, H, W, C, pad, halo
        self.Hp, self.Wp = H + 2 * pad, W + 2 * pad
        self.t = torch.zeros((N, self.Hp, self.Wp, C), device=device, dtype=dtype)

    def _strides(self):
        return self.Hp * self.Wp * self.C, self.Wp * self.C, self.C

    def interior(self):
        sn, sy, sx = self._strides()
        return V(self.t, self.pad * sy + self.pad * sx, self.N, self.H, self.W, self.C, sn, sy, sx)

    def padded(self):
        sn, sy, sx = self._strides()
        return V(self.t, 0, self.N, self.Hp, self.Wp, self.C, sn, sy, sx)


# ------------------------------------------------------------------------------------------------
# operator specs
# ------------------------------------------------------------------------------------------------
class ConvSpec:
    kind = "conv"

    def __init__(self, name):
        self.name = name
        self.src = []          # list[V]
        self.taps = []         # list[(src, dy, dx, c0)]
        self.weight = None     # [rows, ntaps*64]
        self.bias = None       # fp32 [rows]
        self.flat = 0
        self.tile = (8, 16)
        self.n_img = self.out_h = self.out_w = 0
        self.flat_hw = 0
        self.cout = 0
        self.bn = 64
        self.stages = 4
        self.act = nv.ACT_NONE
        self.epi = nv.EPI_STD
        self.out_t = None
        self.out_off = 0
        self.out_fp32 = 0
        self.out_strides = (0, 0, 0)
        self.out_scale, self.out_oy, self.out_ox = 1, 0, 0
        self.halo = nv.HALO_NONE
        self.res = None        # V
        self.res_relu = 0
        self.grouped = 0
        self.out2 = None
        self.n_cls = 0
        self.macs = 0          # algorithmic multiply-accumulates of the reference layer(s) this op computes
        self.group = ""        # subsystem tag for per-op timing
        # row groups (flat mode): stacked problems sharing the weights
        self.group_end = []    # exclusive row end per group
        self.group_scale = None  # fp32 [n_groups, rows] or None
        self.group_shift = None
        self.group_addr = 0
        self.group_hw = []
        self.group_out_base = []

    def to_desc(self):
        d = nv.ConvDesc()
        assert len(self.src) <= nv.HN_MAX_SRC and len(self.taps) <= nv.HN_MAX_TAPS, self.name
        for i, v in enumerate(self.src):
            d.src[i] = v.to_c()
        d.n_src = len(self.src)
        d.weight = self.weight.data_ptr()
        d.w_rows = self.weight.shape[0]
        d.num_taps = len(self.taps)
        assert self.weight.shape[1] == 64 * len(self.taps), self.name
        for i, (s, dy, dx, c0) in enumerate(self.taps):
            d.taps[i] = nv.Tap(s, dy, dx, 0, c0, 0)
        d.flat = self.flat
        d.tile_h, d.tile_w = self.tile
        d.n_img, d.out_h, d.out_w, d.flat_hw = self.n_img, self.out_h, self.out_w, self.flat_hw
        d.cout, d.bn, d.stages = self.cout, self.bn, self.stages
        d.bias = self.bias.data_ptr() if self.bias is not None else None
        d.act, d.epi = self.act, self.epi
        d.out = self.out_t.data_ptr() + self.out_t.element_size() * self.out_off
        d.out_fp32 = self.out_fp32
        d.out_stride_n, d.out_stride_y, d.out_stride_x = self.out_strides
        d.out_scale, d.out_oy, d.out_ox, d.halo = self.out_scale, self.out_oy, self.out_ox, self.halo
        if self.res is not None:
            d.res = self.res.t.data_ptr() + self.res.t.element_size() * self.res.off
            d.res_stride_n, d.res_stride_y, d.res_stride_x = self.res.sn, self.res.sy, self.res.sx
        d.res_relu, d.grouped = self.res_relu, self.grouped
        d.out2 = self.out2.data_ptr() if self.out2 is not None else None
        d.n_cls = self.n_cls
        d.n_groups = len(self.group_end)
        for i, e in enumerate(self.group_end):
            d.group_end[i] = e
        if self.group_shift is not None:
            assert self.group_shift.shape == (len(self.group_end), self.weight.shape[0]), self.name
            d.group_shift = self.group_shift.data_ptr()
            d.group_scale = self.group_scale.data_ptr() if self.group_scale is not None else None
        d.group_addr = self.group_addr
        for i, (hw, base) in enumerate(zip(self.group_hw, self.group_out_base)):
            d.group_hw[i], d.group_out_base[i] = hw, base
        return d

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_conv(plan, self.to_desc()))

    launches = 1


class StemSpec:
    kind, launches, group = "stem", 1, "backbone"

    def __init__(self, x, w, b, out):
        self.x, self.w, self.b, self.out, self.name = x, w, b, out, "stem"
        self.macs = out.N * out.H * out.W * 32 * 27

    def to_desc(self):
        N, _, H, W = self.x.shape
        return nv.StemDesc(self.x.data_ptr(), N, H, W, self.w.data_ptr(), self.b.data_ptr(), self.out.to_c())

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_stem(plan, self.to_desc()))


class NodeSpec:
    kind, launches = "node", 1

    def __init__(self, name, ins, modes, ws, swish, dw, out, group):
        self.name, self.ins, self.modes, self.ws, self.swish, self.dw, self.out, self.group = \
            name, ins, modes, ws, swish, dw, out, group
        self.macs = out.N * out.H * out.W * out.C * 9

    def to_desc(self):
        d = nv.NodeDesc()
        d.n_in = len(self.ins)
        for i, v in enumerate(self.ins):
            d.in_[i] = v.to_c()
            d.mode[i] = self.modes[i]
            d.w[i] = self.ws[i]
        d.swish = self.swish
        d.dw = self.dw.data_ptr()
        d.out = self.out.to_c()
        return d

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_node(plan, self.to_desc()))


class DwMultiSpec:
    kind, launches = "dw_multi", 1

    def __init__(self, name, ins, outs, dw, group):
        self.name, self.ins, self.outs, self.dw, self.group = name, ins, outs, dw, group
        self.macs = sum(o.N * o.H * o.W * o.C * 9 for o in outs)

    def to_desc(self):
        d = nv.DwMultiDesc()
        d.n = len(self.ins)
        for i, (a, b) in enumerate(zip(self.ins, self.outs)):
            d.in_[i], d.out[i] = a.to_c(), b.to_c()
        d.dw = self.dw.data_ptr()
        return d

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_dw_multi(plan, self.to_desc()))


class PoolSpec:
    kind, launches, macs = "pool", 1, 0

    def __init__(self, name, vin, vout, mode, group):
        self.name, self.vin, self.vout, self.mode, self.group = name, vin, vout, mode, group

    def to_desc(self):
        return nv.PoolDesc(self.vin.to_c(), self.vout.to_c(), self.mode)

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_pool(plan, self.to_desc()))


class LaneFuseSpec:
    kind, launches, macs, group = "lanefuse", 1, 0, "lane"

    def __init__(self, p3, p4, p5, p6, out, stride):
        self.p3, self.p4, self.p5, self.p6, self.out, self.stride, self.name = p3, p4, p5, p6, out, stride, "lane.fuse"

    def to_desc(self):
        return nv.LaneFuseDesc(self.p3.to_c(), self.p4.to_c(), self.p5.to_c(), self.p6.to_c(), self.out.to_c(), self.stride)

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_lanefuse(plan, self.to_desc()))


class SePoolSpec:
    """Global average pool and, when FC weights are given, the two squeeze-excite FC layers (one launch)."""
    kind, launches, group, macs = "se_pool", 1, "backbone", 0

    def __init__(self, name, x, pix, partial, counter, mean, fc=None):
        self.name, self.x, self.pix, self.partial, self.counter, self.mean = name, x, pix, partial, counter, mean
        self.fc = fc  # None or dict(S, w1 bf16 [S][C], b1 fp32 [S], w2 bf16 [C][S], b2 fp32 [C], gate bf16 [N][C])
        if fc is not None:
            self.macs = 2 * x.N * fc["S"] * x.C

    def to_desc(self):
        d = nv.SePoolDesc(self.x.to_c(), self.pix, self.partial.data_ptr(), self.counter.data_ptr(), self.mean.data_ptr())
        if self.fc is not None:
            f = self.fc
            d.S, d.w1, d.b1, d.w2, d.b2, d.gate = f["S"], f["w1"].data_ptr(), f["b1"].data_ptr(), f["w2"].data_ptr(), f["b2"].data_ptr(), f["gate"].data_ptr()
        return d

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_se_pool(plan, self.to_desc()))


class SeFusedSpec(SePoolSpec):
    """The whole squeeze-excite of a block in one launch (hn_se_fused_fwd): pool, both FC layers and the in-place scaling."""
    kind = "se_fused"

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_se_fused(plan, self.to_desc()))


class GconvSeSpec(SePoolSpec):
    """An XBlock's grouped 3x3 (stride 1, BN folded, ReLU) and its squeeze-excite in one launch (hn_gconv_se_fwd)."""
    kind = "gconv_se"

    def __init__(self, name, vin, w_ref, b_ref, wq, bias, x, pix, partial, counter, mean, fc):
        super().__init__(name, x, pix, partial, counter, mean, fc)
        self.vin, self.w_ref, self.b_ref, self.wq, self.cbias = vin, w_ref, b_ref, wq, bias
        self.macs += x.N * x.H * x.W * x.C * 8 * 9

    def to_desc(self):
        return nv.GconvSeDesc(self.vin.to_c(), self.wq.data_ptr(), self.cbias.data_ptr(), super().to_desc())

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_gconv_se(plan, self.to_desc()))


class DetPostSpec:
    """Detection decode + NMS (DetectionHeader.decode_device) as an op of the plan: runs in the detection branch, so it
    overlaps the segmentation head instead of following the whole forward."""
    kind, launches, group, macs = "det_post", 12, "detect", 0

    def __init__(self, name, anchors, reg, cls, img_hw, conf_thres, iou_thres, dev):
        self.name = name
        N, A, ncls = cls.shape
        self.anchors, self.reg, self.cls = anchors, reg, cls
        self.ws = torch.zeros(nv.lib.hn_det_workspace_bytes(N, A), dtype=torch.uint8, device=dev)
        self.boxes = torch.zeros((N, A, 4), dtype=torch.float32, device=dev)
        self.scores = torch.zeros((N, A), dtype=torch.float32, device=dev)
        self.cids = torch.zeros((N, A), dtype=torch.int64, device=dev)
        self.count = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.cand = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.desc = nv.DetDesc(anchors.data_ptr(), reg.data_ptr(), cls.data_ptr(), N, A, ncls, int(img_hw[0]), int(img_hw[1]),
                               float(conf_thres), float(iou_thres), int(nv.NMS_AUTO_CUDA), self.ws.data_ptr(), self.ws.numel(),
                               self.boxes.data_ptr(), self.scores.data_ptr(), self.cids.data_ptr(), self.count.data_ptr(),
                               self.cand.data_ptr(), None)

    def result(self):
        return self.boxes, self.scores, self.cids, self.count, self.cand

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_det(plan, self.desc))


class LanePostSpec:
    """Lane decode + lane NMS (LaneHeader.decode_device) as an op of the plan (lane branch)."""
    kind, launches, group, macs = "lane_post", 1, "lane", 0

    def __init__(self, name, pcls, ploc, codec, conf_thres, nms_thres, use_mean, dev):
        import numpy as np
        self.name = name
        N, na = pcls.shape[0], pcls.shape[1]
        fh, fw, ppl = codec.feature_height, codec.feature_width, codec.points_per_line
        assert na == fh * fw and ploc.shape[2] == 2 * ppl + 2, "lane predictions do not match the codec geometry"
        self.ws = torch.empty(nv.lib.hn_lane_workspace_bytes(N, na, ppl), dtype=torch.uint8, device=dev)
        self.count = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.cand = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.meta = torch.zeros((N, na, 4), dtype=torch.int32, device=dev)
        self.prob = torch.zeros((N, na), dtype=torch.float32, device=dev)
        self.xs = torch.zeros((N, na, ppl), dtype=torch.float32, device=dev)
        self.desc = nv.LaneDesc(pcls.data_ptr(), ploc.data_ptr(), N, fh, fw, ppl, 0, float(np.float32(conf_thres)),
                                float(np.float32(nms_thres)), int(bool(use_mean)), float(codec.step_w), float(codec.interval),
                                float(codec.points_per_anchor), float(codec.input_width), 100.0, self.ws.data_ptr(),
                                self.count.data_ptr(), self.meta.data_ptr(), self.prob.data_ptr(), self.xs.data_ptr(),
                                self.cand.data_ptr())

    def result(self):
        return self.count, self.meta, self.prob, self.xs, self.cand

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_lane(plan, self.desc))


class WaitSpec:
    """Branch dependency (hn_plan_add_wait): the ops after it in this branch start once branch ``wait_for`` has finished."""
    kind, launches, macs = "wait", 0, 0

    def __init__(self, name, wait_for, group):
        self.name, self.wait_for, self.group = name, wait_for, group

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_wait(plan, self.wait_for))


class SeScaleSpec:
    kind, launches, group, macs = "se_scale", 1, "backbone", 0

    def __init__(self, name, x, scale):
        self.name, self.x, self.scale = name, x, scale

    def to_desc(self):
        return nv.SeScaleDesc(self.x.to_c(), self.scale.data_ptr())

    def add_to(self, plan):
        nv.check(nv.lib.hn_plan_add_se_scale(plan, self.to_desc()))


# ------------------------------------------------------------------------------------------------
# weight folding / packing helpers
# ------------------------------------------------------------------------------------------------
def fold_bn(weight, bias, bn):
    """conv(+bias) followed by eval-mode BatchNorm -> equivalent (weight, bias), fp32."""
    w = weight.detach().float()
    b = bias.detach().float() if bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w, b


def choose_bn(cout):
    """N tile: a multiple of 16, at most 256; layers wider than one tile use multiples of 64 (the TMA-store slab)."""
    if cout <= 256:
        return max(16, (cout + 15) // 16 * 16)
    n_tiles = (cout + 255) // 256
    per = (cout + n_tiles - 1) // n_tiles
    return (per + 63) // 64 * 64


def choose_stages(bn, ntaps):
    """Shared-memory ring depth: leave room for the output staging slabs (16 KB each) and, for narrow N tiles,
    for a second co-resident CTA (227 KB and 512 TMEM columns per SM)."""
    stage = 16384 + bn * 128
    budget = (227 - 16 - 5) * 1024 if bn > 128 else (113 - 32 - 5) * 1024
    # the ring runs across the tiles of a persistent CTA, so its depth is not capped by the taps of one tile: short-K
    # layers are bound by bytes in flight (load latency x ring depth)
    return int(max(2, min(8, budget // stage)))


def choose_tile(H, W):
    best = None
    for th, tw in ((8, 16), (16, 8), (4, 32), (32, 4), (2, 64), (1, 128), (64, 2)):
        n = math.ceil(H / th) * math.ceil(W / tw)
        if best is None or n < best[0]:
            best = (n, (th, tw))
    return best[1]


def pack_weight(entries, cout, bn, wdtype, device):
    """entries: [(src, dy, dx, W[cout, Csrc])] -> (taps, K-major matrix [rows_pad, ntaps*64]).  Assembled where the weights
    live (the CPU shadow of the parameters) and uploaded once."""
    taps, spans, k = [], [], 0
    for (s, dy, dx, wt) in entries:
        cs = wt.shape[1]
        for c0 in range(0, cs, 64):
            taps.append((s, dy, dx, c0))
            spans.append((wt, c0, min(64, cs - c0), k))
            k += 64
    rows = (cout + bn - 1) // bn * bn
    out = torch.zeros((rows, k), dtype=torch.float32, device=entries[0][3].device)
    for wt, c0, w, k0 in spans:
        out[:cout, k0:k0 + w] = wt[:, c0:c0 + w]
    return taps, out.to(wdtype).contiguous().to(device)


def pad_bias(b, cout, bn, device=None):
    rows = (cout + bn - 1) // bn * bn
    out = torch.zeros(rows, dtype=torch.float32, device=b.device)
    out[:cout] = b
    return out.to(device) if device is not None else out


# row/column tap sets of the parity-collapsed up-sampled 3x3 conv: parity -> [(source shift, [k...])]
UP_SETS = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}


class Builder:
    """Emits the op list for one (batch, height, width)."""

    def __init__(self, model, B, H, W, device, act_dtype=torch.bfloat16, out_alloc=None):
        self.m, self.B, self.H, self.W, self.dev, self.dt = model, B, H, W, device, act_dtype
        self.ops = []
        self.keep = []  # tensors that must stay alive
        self.out = {}
        self.out_alloc = out_alloc  # optional (name, shape, dtype) -> tensor: lets a caller own the head outputs

    def new_out(self, name, shape, dtype):
        """A head output tensor [B, ...]: fresh zeros, or the caller's buffer (e.g. a batch slice of a larger one)."""
        if self.out_alloc is not None:
            t = self.out_alloc(name, tuple(shape), dtype)
            assert tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.is_contiguous(), name
            return t
        return torch.zeros(tuple(shape), dtype=dtype, device=self.dev)

    # -- small helpers --
    def buf(self, H, W, C, pad=0, halo=nv.HALO_NONE):
        return Buf(self.dev, self.dt, self.B, H, W, C, pad, halo)

    def f32(self, t):
        t = t.detach().float().contiguous().to(self.dev)
        self.keep.append(t)
        return t

    #: layers whose weights are packed as a bf16 (hi, lo) PAIR: W = hi + lo with hi = bf16(W), lo = bf16(W - hi), two K blocks per
    #: original block reading the same activations -- the GEMM then sees the weights to ~2^-17 instead of 2^-9.  Measured
    #: on the B200 at 640x640 over three weight seeds: regression error 0.0090 / 0.0047 / 0.0070 -> 0.0066 / 0.0039 / 0.0048 of max|ref|,
    #: lane logits 0.0124 -> 0.0085 on the worst seed.  The detection / lane GEMMs are latency-bound (K doubles for free: 8.42 ms
    #: per step either way).  The same trick on the seg decoder's last layers buys 0.02-0.12 % of arg-max agreement but costs
    #: 0.3-0.6 ms (HN_HILO=det.,lane.,seg.out[,seg.d7,seg.d6]: 8.72 / 9.02 ms), so it stays off.
    HILO_PREFIXES = tuple(p for p in os.environ.get("HN_HILO", "det.,lane.").split(",") if p and p != "none")

    def _finish(self, cs, entries, cout, bias, bn=None):
        if self.dt == torch.bfloat16 and cs.name.startswith(tuple(getattr(self.m, "hilo_prefixes", self.HILO_PREFIXES))):
            split = []
            for (s_, dy, dx, wt) in entries:
                hi = wt.to(torch.bfloat16).float()
                split += [(s_, dy, dx, hi), (s_, dy, dx, wt - hi)]
            entries = split
        cs.cout = cout
        cs.bn = bn or choose_bn(cout)
        cs.taps, cs.weight = pack_weight(entries, cout, cs.bn, self.dt, self.dev)
        cs.bias = pad_bias(bias, cout, cs.bn, self.dev)
        cs.stages = choose_stages(cs.bn, len(cs.taps))
        self.ops.append(cs)
        return cs

    def _set_out_buf(self, cs, ob, vin_like=None):
        """Route the output into buffer ``ob`` (interior), choosing flat or spatial tiling."""
        vi = ob.interior()
        cs.out_t, cs.out_off = ob.t, vi.off
        cs.out_strides = (vi.sn, vi.sy, vi.sx)
        cs.halo = ob.halo if ob.pad else nv.HALO_NONE
        cs.n_img, cs.out_h, cs.out_w = ob.N, ob.H, ob.W

    def conv1x1(self, name, vin, w, b, ob, act, group, res=None, res_relu=0, out_chan=None):
        """1x1 conv of view ``vin`` into buffer ``ob`` (optionally a channel slice)."""
        cs = ConvSpec(name)
        cs.group, cs.act, cs.res, cs.res_relu = group, act, res, res_relu
        cout = w.shape[0]
        self._set_out_buf(cs, ob)
        flat = vin.is_flattenable() and ob.pad == 0 and (res is None or res.is_flattenable())
        if flat:
            cs.flat, cs.flat_hw = 1, vin.H * vin.W
            cs.src = [vin.flat()]
            cs.out_strides = (ob.H * ob.W * ob.C, 0, ob.C)
            if res is not None:
                cs.res = V(res.t, res.off, 1, 1, res.N * res.H * res.W, res.C, res.H * res.W * res.sx, 0, res.sx)
        else:
            cs.src = [vin]
            cs.tile = choose_tile(ob.H, ob.W)
        cs.macs = ob.N * ob.H * ob.W * cout * w.shape[1]
        return self._finish(cs, [(0, 0, 0, w.reshape(cout, -1))], cout, b)

    # -- backbone --
    def backbone(self, x_static):
        net = self.m.backbone.net
        B = self.B
        H1, W1 = (self.H + 1) // 2, (self.W + 1) // 2
        stem = net.stem
        w, b = fold_bn(stem.conv.weight, None, stem.bn)
        wk = self.f32(w.permute(1, 2, 3, 0).reshape(27, 32))
        sb = self.buf(H1, W1, 32)
        self.ops.append(StemSpec(x_static, wk, self.f32(b), sb.interior()))
        cur, curH, curW = sb, H1, W1
        feats = []
        n_stage = self.m.backbone.stage_num
        seg_on = self.m.segheader is not None
        for s in range(n_stage):
            stage = getattr(net, "stage_%d" % s)
            blocks = list(stage.blocks.children())
            for bi, blk in enumerate(blocks):
                nm = "backbone.s%d.b%d" % (s, bi)
                st, mid, cout = blk.stride, blk.mid, blk.cout
                Ho, Wo = (curH - 1) // st + 1, (curW - 1) // st + 1
                vin = cur.interior()
                # 1x1 + BN + ReLU
                w1, b1 = fold_bn(blk.conv_block_1[0].weight, None, blk.conv_block_1[1])
                a = self.buf(curH, curW, mid)
                self.conv1x1(nm + ".c1", vin, w1, b1, a, nv.ACT_RELU, "backbone")
                # grouped 3x3 (+BN+ReLU), stride st
                g = self.buf(Ho, Wo, mid)
                # small maps, stride 1: the grouped 3x3 and the whole squeeze-excite are ONE launch
                if not (st == 1 and blk.se is not None and self.squeeze_excite(nm + ".se", g, blk.se, conv=(a, blk))):
                    self.gconv(nm + ".c2", a, blk, g, st)
                    # squeeze-excite (in place): pool -> FC1+ReLU -> FC2+sigmoid -> scale
                    if blk.se is not None:
                        self.squeeze_excite(nm + ".se", g, blk.se)
                # shortcut
                if blk.shortcut is not None:
                    ws, bs = fold_bn(blk.shortcut[0].weight, None, blk.shortcut[1])
                    sc = self.buf(Ho, Wo, cout)
                    vs = vin.phase(0, 0) if st == 2 else vin
                    self.conv1x1(nm + ".sc", vs, ws, bs, sc, nv.ACT_NONE, "backbone")
                    res = sc.interior()
                else:
                    res = vin
                last = bi == len(blocks) - 1
                pad = 1 if (last and s == 0 and seg_on) else 0
                ob = self.buf(Ho, Wo, cout, pad, nv.HALO_REFLECT if pad else nv.HALO_NONE)
                w3, b3 = fold_bn(blk.conv_block_3[0].weight, None, blk.conv_block_3[1])
                self.conv1x1(nm + ".c3", g.interior(), w3, b3, ob, nv.ACT_NONE, "backbone", res=res, res_relu=1)
                cur, curH, curW = ob, Ho, Wo
            feats.append(cur)
        return feats

    def fc(self, name, vin_rows, w, b, out_t, act):
        """rows x C matrix (a [1,1,rows,C] view) times w[cout, C] on the tensor cores, bf16 out_t [rows, cout]."""
        cs = ConvSpec(name)
        cs.group, cs.act = "backbone", act
        cs.flat, cs.flat_hw = 1, vin_rows.W
        cs.src = [vin_rows]
        cs.out_t, cs.out_off = out_t, 0
        cout = w.shape[0]
        cs.out_strides = (vin_rows.W * cout, 0, cout)
        cs.macs = vin_rows.W * cout * w.shape[1]
        # one M tile only (rows = batch): narrow N tiles spread the weight read over many SMs
        return self._finish(cs, [(0, 0, 0, w)], cout, b, bn=16 if cout <= 256 else 64)

    def squeeze_excite(self, name, g, se, conv=None):
        """``conv`` = (input Buf, block): also run the block's stride-1 grouped 3x3 in the same launch when the shape allows;
        returns True when it did (the caller then skips the separate conv op)."""
        B, C, S = self.B, g.C, se[1].weight.shape[0]
        Sp = (S + 7) // 8 * 8  # hidden width padded to the 16-byte channel granule
        dev = self.dev
        hw = g.H * g.W
        pix = min(2048, max(128, (hw // 16 + 127) // 128 * 128))  # ~16 blocks per image, 128..2048 pixels each
        partial = torch.zeros((B, (hw + pix - 1) // pix, C), dtype=torch.float32, device=dev)
        counter = torch.zeros((B,), dtype=torch.int32, device=dev)
        mean = torch.zeros((B, C), dtype=self.dt, device=dev)
        scale = torch.zeros((B, C), dtype=self.dt, device=dev)
        w1 = torch.zeros((Sp, C), dtype=torch.float32, device=se[1].weight.device)
        b1 = torch.zeros((Sp,), dtype=torch.float32, device=w1.device)
        w1[:S], b1[:S] = se[1].weight.detach().float().reshape(S, C), se[1].bias.detach().float()
        w2 = torch.zeros((C, Sp), dtype=torch.float32, device=w1.device)
        w2[:, :S] = se[3].weight.detach().float().reshape(C, S)
        # pool -> FC1 + ReLU -> FC2 + sigmoid in ONE launch: the block that finishes an image's pooling runs its FCs
        fc = dict(S=Sp, w1=w1.to(self.dt).contiguous().to(dev), b1=b1.to(dev), w2=w2.to(self.dt).contiguous().to(dev),
                  b2=se[3].bias.detach().float().contiguous().to(dev), gate=scale)
        if conv is not None and GCONV_SE and nv.lib.hn_gconv_se_supported(g.H, g.W, C, Sp):
            a, blk = conv
            w, b = fold_bn(blk.conv_block_2[0].weight, None, blk.conv_block_2[1])  # [C, 8, 3, 3]
            assert w.shape[1] == 8 and C % 8 == 0, "grouped conv path assumes group width 8"
            wq = torch.zeros((C // 8, 10, 8, 8), dtype=torch.float32, device=w.device)  # [group][tap][out][in], 10th tap zero
            wq[:, :9] = w.detach().float().reshape(C // 8, 8, 8, 9).permute(0, 3, 1, 2)
            w_ref = w.detach().float().to(self.dt).float().contiguous().to(dev)
            self.ops.append(GconvSeSpec(name[:-3] + ".c2se", a.interior(), w_ref, b.detach().float().to(dev), wq.to(self.dt).contiguous().to(dev),
                                        self.f32(b), g.interior(), pix, partial, counter, mean, fc))
            return True
        if conv is not None:
            return False
        if SE_FUSED and nv.lib.hn_se_fused_supported(g.H, g.W, C, Sp):  # small maps: one cluster launch per block
            self.ops.append(SeFusedSpec(name, g.interior(), pix, partial, counter, mean, fc))
            return
        self.ops.append(SePoolSpec(name + ".gate", g.interior(), pix, partial, counter, mean, fc))
        self.ops.append(SeScaleSpec(name + ".scale", g.interior(), scale))

    def gconv(self, name, a, blk, ob, stride):
        conv, bn = blk.conv_block_2[0], blk.conv_block_2[1]
        w, b = fold_bn(conv.weight, None, bn)  # [C, gw, 3, 3]
        C, gw = w.shape[0], w.shape[1]
        assert gw == 8 and C % 8 == 0, "grouped conv path assumes group width 8"
        cs = ConvSpec(name)
        cs.group, cs.act, cs.grouped = "backbone", nv.ACT_RELU, 1
        self._set_out_buf(cs, ob)
        cs.tile = choose_tile(ob.H, ob.W)
        va = a.interior()
        if stride == 1:
            cs.src = [va]
            where = {(ky, kx): (0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)}
        else:
            cs.src = [va.phase(0, 0), va.phase(0, 1), va.phase(1, 0), va.phase(1, 1)]
            # input row 2y+ky-1 = 2(y+ay)+ry
            ph = {0: (1, -1), 1: (0, 0), 2: (1, 0)}
            where = {(ky, kx): (ph[ky][0] * 2 + ph[kx][0], ph[ky][1], ph[kx][1]) for ky in range(3) for kx in range(3)}
        bn_t = 64
        rows = (C + bn_t - 1) // bn_t * bn_t
        wm = torch.zeros((rows, 9 * 64), dtype=torch.float32, device=w.device)
        co = torch.arange(C, device=w.device)
        base = (co // 8) * 8 - (co // 64) * 64
        taps = []
        for ky in range(3):
            for kx in range(3):
                t = ky * 3 + kx
                s, dy, dx = where[(ky, kx)]
                taps.append((s, dy, dx, 0))
                for j in range(8):
                    wm[co, t * 64 + base + j] = w[:, j, ky, kx]
        cs.cout, cs.bn = C, bn_t
        cs.taps, cs.weight = taps, wm.to(self.dt).contiguous().to(self.dev)
        cs.bias = pad_bias(b, C, bn_t, self.dev)
        cs.stages = choose_stages(bn_t, 9)
        cs.macs = ob.N * ob.H * ob.W * C * 8 * 9
        self.ops.append(cs)

    # -- neck --
    def sepconv(self, name, ins, modes, ws, swish, sep, bn, ob, act, group, h, w, out_fp32=None):
        """node/depthwise kernel -> pointwise GEMM (+bias, folded BN, activation)."""
        C = sep.depthwise_conv.conv.weight.shape[0]
        tmp = self.buf(h, w, C)
        dw = self.f32(sep.depthwise_conv.conv.weight.reshape(C, 9).t())
        self.ops.append(NodeSpec(name + ".dw", ins, modes, ws, swish, dw, tmp.interior(), group))
        pw, pb = fold_bn(sep.pointwise_conv.conv.weight, sep.pointwise_conv.conv.bias, bn)
        if out_fp32 is None:
            return self.conv1x1(name + ".pw", tmp.interior(), pw, pb, ob, act, group)
        # fp32 head output: rows of a [B, total, k] tensor
        t, row0, k = out_fp32
        cs = ConvSpec(name + ".pw")
        cs.group, cs.act = group, act
        cs.flat, cs.flat_hw = 1, h * w
        cs.src = [tmp.interior().flat()]
        cs.out_t, cs.out_off, cs.out_fp32 = t, row0 * k, 1
        cout = pw.shape[0]
        cs.out_strides = (t.shape[1] * k, 0, cout)
        cs.macs = self.B * h * w * cout * C
        return self._finish(cs, [(0, 0, 0, pw.reshape(cout, -1))], cout, pb)

    def neck(self, feats):
        cells = list(self.m.neck.bifpn.children())
        seg_on = self.m.segheader is not None
        levels = None
        for ci, cell in enumerate(cells):
            nm = "neck.c%d" % ci
            last = ci == len(cells) - 1
            eps = cell.epsilon
            if cell.first_time:
                if len(feats) == 4:
                    c3, c4, c5 = feats[-3:]
                    r = cell.p5_to_p6
                    w, b = fold_bn(r[0].conv.weight, r[0].conv.bias, r[1])
                    t6 = self.buf(c5.H, c5.W, w.shape[0])
                    self.conv1x1(nm + ".p5_to_p6", c5.interior(), w, b, t6, nv.ACT_NONE, "neck")
                    p6_in = self.buf((c5.H - 2) // 2 + 1, (c5.W - 2) // 2 + 1, w.shape[0])
                    self.ops.append(PoolSpec(nm + ".p6pool", t6.interior(), p6_in.interior(), nv.POOL_ZERO_RB, "neck"))
                else:
                    c3, c4, c5, c6 = feats[-4:]
                    r = cell.p6_down_channel
                    w, b = fold_bn(r[0].conv.weight, r[0].conv.bias, r[1])
                    p6_in = self.buf(c6.H, c6.W, w.shape[0])
                    self.conv1x1(nm + ".p6_down", c6.interior(), w, b, p6_in, nv.ACT_NONE, "neck")
                ch = p6_in.C
                p7_in = self.buf((p6_in.H - 2) // 2 + 1, (p6_in.W - 2) // 2 + 1, ch)
                self.ops.append(PoolSpec(nm + ".p7pool", p6_in.interior(), p7_in.interior(), nv.POOL_ZERO_RB, "neck"))
                r = cell.p3_down_channel
                w, b = fold_bn(r[0].conv.weight, r[0].conv.bias, r[1])
                p3b = self.buf(c3.H, c3.W, ch)
                self.conv1x1(nm + ".p3_down", c3.interior(), w, b, p3b, nv.ACT_NONE, "neck")

                def two(ra, rb, c, tag):  # the two reducers of one level share their input: one GEMM
                    wa, ba = fold_bn(ra[0].conv.weight, ra[0].conv.bias, ra[1])
                    wb, bb = fold_bn(rb[0].conv.weight, rb[0].conv.bias, rb[1])
                    ob = self.buf(c.H, c.W, 2 * ch)
                    self.conv1x1(nm + tag, c.interior(), torch.cat([wa, wb], 0), torch.cat([ba, bb], 0), ob, nv.ACT_NONE, "neck")
                    return ob.interior().chan(0, ch), ob.interior().chan(ch, ch)

                p4_a, p4_b = two(cell.p4_down_channel, cell.p4_down_channel_2, c4, ".p4_down")
                p5_a, p5_b = two(cell.p5_down_channel, cell.p5_down_channel_2, c5, ".p5_down")
                p3_in, p6v, p7v = p3b.interior(), p6_in.interior(), p7_in.interior()
            else:
                p3_in, p4_a, p5_a, p6v, p7v = [l.interior() for l in levels]
                p4_b, p5_b = p4_a, p5_a
                ch = p3_in.C

            def wts(p):
                w = torch.relu(p.detach().float())
                w = w / (torch.sum(w, dim=0) + eps)
                return [float(v) for v in w.cpu()]

            def node(tag, sep, ins, modes, ws, h, w, pad=0):
                ob = self.buf(h, w, ch, pad, nv.HALO_REFLECT if pad else nv.HALO_NONE)
                self.sepconv(nm + "." + tag, ins, modes, ws, 1, sep, sep.bn, ob, nv.ACT_NONE, "neck", h, w)
                return ob

            S, U, P = nv.IN_SAME, nv.IN_UP2, nv.IN_POOL
            p6_up = node("conv6_up", cell.conv6_up, [p6v, p7v], [S, U], wts(cell.p6_w1), p6v.H, p6v.W)
            p5_up = node("conv5_up", cell.conv5_up, [p5_a, p6_up.interior()], [S, U], wts(cell.p5_w1), p5_a.H, p5_a.W)
            p4_up = node("conv4_up", cell.conv4_up, [p4_a, p5_up.interior()], [S, U], wts(cell.p4_w1), p4_a.H, p4_a.W)
            segpad = 1 if (last and seg_on) else 0
            p3_out = node("conv3_up", cell.conv3_up, [p3_in, p4_up.interior()], [S, U], wts(cell.p3_w1), p3_in.H, p3_in.W, segpad)
            p4_out = node("conv4_down", cell.conv4_down, [p4_b, p4_up.interior(), p3_out.interior()], [S, S, P],
                          wts(cell.p4_w2), p4_a.H, p4_a.W, segpad)
            p5_out = node("conv5_down", cell.conv5_down, [p5_b, p5_up.interior(), p4_out.interior()], [S, S, P],
                          wts(cell.p5_w2), p5_a.H, p5_a.W, segpad)
            p6_out = node("conv6_down", cell.conv6_down, [p6v, p6_up.interior(), p5_out.interior()], [S, S, P],
                          wts(cell.p6_w2), p6v.H, p6v.W)
            p7_out = node("conv7_down", cell.conv7_down, [p7v, p6_out.interior()], [S, P], wts(cell.p7_w2), p7v.H, p7v.W)
            levels = [p3_out, p4_out, p5_out, p6_out, p7_out]
        return levels

    # -- segmentation head --
    def conv3x3_plain(self, name, ib, w, b, ob, act):
        """3x3 over a reflect-padded buffer."""
        cs = ConvSpec(name)
        cs.group, cs.act = "seg", act
        self._set_out_buf(cs, ob)
        cs.tile = choose_tile(ob.H, ob.W)
        cs.src = [ib.padded()]
        cout = w.shape[0]
        entries = [(0, ky, kx, w[:, :, ky, kx]) for ky in range(3) for kx in range(3)]
        cs.macs = ob.N * ob.H * ob.W * cout * w.shape[1] * 9
        return self._finish(cs, entries, cout, b)

    def conv3x3_up(self, name, upb, skipb, w, b, ob, act):
        """3x3 (reflect pad) over cat(up2(upb), skipb) as four parity-collapsed convs."""
        cout, cup = w.shape[0], upb.C
        assert upb.pad == 1 and upb.halo == nv.HALO_REPLICATE
        for py in range(2):
            for px in range(2):
                cs = ConvSpec("%s.p%d%d" % (name, py, px))
                cs.group, cs.act = "seg", act
                self._set_out_buf(cs, ob)
                # tile space = source resolution, output pixel (2y+py, 2x+px)
                cs.out_h, cs.out_w = upb.H, upb.W
                cs.out_scale, cs.out_oy, cs.out_ox = 2, py, px
                cs.tile = choose_tile(upb.H, upb.W)
                cs.src = [upb.padded()]
                entries = []
                for (sy, kys) in UP_SETS[py]:
                    for (sx, kxs) in UP_SETS[px]:
                        wt = sum(w[:, :cup, ky, kx] for ky in kys for kx in kxs)
                        entries.append((0, sy + 1, sx + 1, wt))
                if skipb is not None:
                    assert skipb.pad == 1 and skipb.halo == nv.HALO_REFLECT
                    sp = skipb.padded()
                    cs.src += [sp.phase(0, 0), sp.phase(0, 1), sp.phase(1, 0), sp.phase(1, 1)]
                    for ky in range(3):
                        for kx in range(3):
                            ry, ay = (py + ky) % 2, (py + ky) // 2
                            rx, ax = (px + kx) % 2, (px + kx) // 2
                            entries.append((1 + ry * 2 + rx, ay, ax, w[:, cup:, ky, kx]))
                cs.macs = (ob.N * ob.H * ob.W * cout * w.shape[1] * 9) // 4
                self._finish(cs, entries, cout, b)

    def seg_out(self, name, ib, w, b, logits, cls_map):
        """final 3x3 over up2(ib): sub-pixel conv at source resolution, 4 parities x 8 columns."""
        ncls = w.shape[0]
        assert ncls <= 8 and ib.pad == 1 and ib.halo == nv.HALO_REPLICATE
        cs = ConvSpec(name)
        cs.group, cs.epi, cs.n_cls = "seg", nv.EPI_SEGOUT, ncls
        cs.src = [ib.padded()]
        cs.n_img, cs.out_h, cs.out_w = ib.N, ib.H, ib.W
        cs.tile = choose_tile(ib.H, ib.W)
        cs.out_t, cs.out_off, cs.out_fp32 = logits, 0, 1
        cs.out2 = cls_map
        entries = []
        bias = torch.zeros(32, dtype=torch.float32, device=w.device)
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                wt = torch.zeros((32, w.shape[1]), dtype=torch.float32, device=w.device)
                for py in range(2):
                    for px in range(2):
                        kys = [k for (s, ks) in UP_SETS[py] if s == sy for k in ks]
                        kxs = [k for (s, ks) in UP_SETS[px] if s == sx for k in ks]
                        r0 = (py * 2 + px) * 8
                        for ky in kys:
                            for kx in kxs:
                                wt[r0:r0 + ncls] += w[:, :, ky, kx]
                entries.append((0, sy + 1, sx + 1, wt))
        for p in range(4):
            bias[p * 8:p * 8 + ncls] = b
        cs.macs = ib.N * ib.H * ib.W * 4 * ncls * w.shape[1] * 9
        return self._finish(cs, entries, 32, bias, bn=32)

    def seg_head(self, feats0, levels):
        dec = list(self.m.segheader.decoder.children())
        n = len(self.m.segheader.num_ch_enc)
        skips = [feats0] + levels[:n - 1]          # input_features
        x = skips[-1]                              # reflect-padded P5 (or deepest feature)
        R, RP = nv.HALO_REFLECT, nv.HALO_REPLICATE
        for i in range(n):
            c0, c1 = dec[2 * i].conv.conv, dec[2 * i + 1].conv.conv
            a = self.buf(x.H, x.W, c0.weight.shape[0], 1, RP)
            self.conv3x3_plain("seg.d%d" % (2 * i), x, c0.weight.detach().float(), c0.bias.detach().float(), a, nv.ACT_ELU)
            skip = skips[n - 2 - i] if i < n - 1 else None
            last = i == n - 1
            ob = self.buf(a.H * 2, a.W * 2, c1.weight.shape[0], 1, RP if last else R)
            self.conv3x3_up("seg.d%d" % (2 * i + 1), a, skip, c1.weight.detach().float(), c1.bias.detach().float(), ob, nv.ACT_ELU)
            x = ob
        oc = dec[-1].conv
        ncls = oc.weight.shape[0]
        logits = self.new_out("seg", (self.B, ncls, x.H * 2, x.W * 2), torch.float32)
        cls_map = self.new_out("seg_cls_u8", (self.B, x.H * 2, x.W * 2), torch.uint8)
        self.seg_out("seg.out", x, oc.weight.detach().float(), oc.bias.detach().float(), logits, cls_map)
        self.out["seg"] = logits
        self.out["seg_cls_u8"] = cls_map

    # -- detection head --
    def det_head(self, levels):
        """Both towers run every layer ONCE over all pyramid levels: the levels' pixels are stacked into one
        [sum_l B*H_l*W_l, C] matrix (level-major), the shared depthwise / pointwise weights are applied in one
        launch each, and the per-level BatchNorm (detection.py:20-24,30-37) becomes a per-row-group scale / shift
        in the GEMM epilogue.  The header GEMMs scatter their rows straight into the reference's
        [B, sum_l H_l*W_l*9, k] layout."""
        dh = self.m.detectheader
        na, ncls = dh.num_anchors, dh.num_classes
        B, C = self.B, levels[0].C
        hws = [l.H * l.W for l in levels]
        total = sum(hws) * na
        reg = self.new_out("regression", (B, total, 4), torch.float32)
        cls = self.new_out("classification", (B, total, ncls), torch.float32)
        ends, acc = [], 0
        for hw in hws:
            acc += B * hw
            ends.append(acc)
        R = acc

        def stacked():
            t = torch.zeros((R, C), dtype=self.dt, device=self.dev)
            views, r0 = [], 0
            for l in levels:
                views.append(V(t, r0 * C, B, l.H, l.W, C, l.H * l.W * C, l.W * C, C))
                r0 += B * l.H * l.W
            return t, views

        def rows_view(t):
            return V(t, 0, 1, 1, R, C, 0, 0, C)

        for tname, tower, out_t, k, act in (("reg", dh.regressor, reg, 4, nv.ACT_NONE), ("cls", dh.classifier, cls, ncls, nv.ACT_SIGMOID)):
            cur = [l.interior() for l in levels]
            for i in range(tower.num_layers):
                sep = tower.conv_list[i]
                dwt, dwv = stacked()
                self.ops.append(DwMultiSpec("det.%s.%d.dw" % (tname, i), cur, dwv,
                                            self.f32(sep.depthwise_conv.conv.weight.reshape(C, 9).t()), "detect"))
                pw = sep.pointwise_conv.conv
                ot, ov = stacked()
                cs = ConvSpec("det.%s.%d.pw" % (tname, i))
                cs.group, cs.act = "detect", nv.ACT_SWISH
                cs.flat, cs.flat_hw = 1, R
                cs.src = [rows_view(dwt)]
                cs.out_t, cs.out_off, cs.out_strides = ot, 0, (R * C, 0, C)
                cs.macs = R * C * C
                cs.group_end = list(ends)
                self._finish(cs, [(0, 0, 0, pw.weight.detach().float().reshape(C, C))], C, torch.zeros(C, device=pw.weight.device))
                scales, shifts = [], []
                for li in range(len(levels)):
                    bn = tower.bn_list[li][i]
                    sc = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
                    sh = (pw.bias.detach().float() - bn.running_mean.detach().float()) * sc + bn.bias.detach().float()
                    scales.append(pad_bias(sc, C, cs.bn, self.dev))
                    shifts.append(pad_bias(sh, C, cs.bn, self.dev))
                cs.group_scale, cs.group_shift = torch.stack(scales).contiguous(), torch.stack(shifts).contiguous()
                cur = ov
            sep = tower.header
            dwt, dwv = stacked()
            self.ops.append(DwMultiSpec("det.%s.hdr.dw" % tname, cur, dwv,
                                        self.f32(sep.depthwise_conv.conv.weight.reshape(C, 9).t()), "detect"))
            pw = sep.pointwise_conv.conv
            cout = pw.weight.shape[0]
            cs = ConvSpec("det.%s.hdr.pw" % tname)
            cs.group, cs.act = "detect", act
            cs.flat, cs.flat_hw = 1, R
            cs.src = [rows_view(dwt)]
            cs.out_t, cs.out_off, cs.out_fp32 = out_t, 0, 1
            cs.out_strides = (total * k, 0, cout)
            cs.macs = R * cout * C
            cs.group_end, cs.group_addr = list(ends), 1
            cs.group_hw = list(hws)
            bases, a0 = [], 0
            for hw in hws:
                bases.append(a0 * k)
                a0 += hw * na
            cs.group_out_base = bases
            self._finish(cs, [(0, 0, 0, pw.weight.detach().float().reshape(cout, C))], cout, pw.bias.detach().float())
        self.out["regression"], self.out["classification"] = reg, cls
        post = getattr(self.m, "_fused_post", None)
        if post and post.get("det") and self.dev.type == "cuda":
            from .heads import make_anchors
            conf, iou = post["det"]
            a = self.m.detectheader.anchors
            anc = torch.from_numpy(make_anchors((self.H, self.W), a.anchor_scale, a.strides, a.scales, a.ratios)).to(self.dev).view(-1, 4).contiguous()
            self.keep.append(anc)
            dp = DetPostSpec("det.post", anc, reg, cls, (self.H, self.W), conf, iou, self.dev)
            self.ops.append(WaitSpec("det.join", 2, "detect"))  # the decode needs the regression tower (branch 2) too
            self.ops.append(dp)
            self.out["det_post"] = dp

    # -- lane head --
    def lane_head(self, levels):
        lh = self.m.laneheader
        p3, p4, p5, p6 = levels[:4]
        if lh.stride == 32:
            fh, fw = p5.H, p5.W
        elif lh.stride == 16:
            fh, fw = p4.H, p4.W
        else:
            raise ValueError("lane anchor_stride must be 16 or 32 (lanedetect.py:70-83)")
        C = p3.C
        fused = self.buf(fh, fw, 4 * C)
        self.ops.append(LaneFuseSpec(p3.interior(), p4.interior(), p5.interior(), p6.interior(), fused.interior(), lh.stride))
        branches = [lh.conv_cls_conv, lh.conv_up_conv, lh.conv_down_conv]
        ws, bs = zip(*[fold_bn(br[0].weight, None, br[1]) for br in branches])
        hid = self.buf(fh, fw, 4 * C * 3)
        self.conv1x1("lane.hidden", fused.interior(), torch.cat(ws, 0), torch.cat(bs, 0), hid, nv.ACT_RELU, "lane")
        ncls = lh.num_classes
        nup, ndown = lh.lane_up_pts_num, lh.lane_down_pts_num
        pcls = self.new_out("predict_cls", (self.B, fh * fw, ncls), torch.float32)
        ploc = self.new_out("predict_loc", (self.B, fh * fw, nup + ndown), torch.float32)
        for bi, (br, t, col0) in enumerate(((lh.conv_cls_conv, pcls, 0), (lh.conv_up_conv, ploc, ndown), (lh.conv_down_conv, ploc, 0))):
            conv = br[3]
            cout = conv.weight.shape[0]
            cs = ConvSpec("lane.out%d" % bi)
            cs.group = "lane"
            cs.flat, cs.flat_hw = 1, fh * fw
            cs.src = [hid.interior().chan(bi * 4 * C, 4 * C).flat()]
            cs.out_t, cs.out_off, cs.out_fp32 = t, col0, 1
            cs.out_strides = (t.shape[1] * t.shape[2], 0, t.shape[2])
            cs.macs = self.B * fh * fw * cout * 4 * C
            self._finish(cs, [(0, 0, 0, conv.weight.detach().float().reshape(cout, -1))], cout, conv.bias.detach().float())
        self.out["predict_cls"], self.out["predict_loc"] = pcls, ploc
        post = getattr(self.m, "_fused_post", None)
        if post and post.get("lane") and self.dev.type == "cuda":
            codec, conf, nms_thres, use_mean = post["lane"]
            lp = LanePostSpec("lane.post", pcls, ploc, codec, conf, nms_thres, use_mean, self.dev)
            self.ops.append(lp)
            self.out["lane_post"] = lp

    def build(self, x_static):
        feats = self.backbone(x_static)
        levels = self.neck(feats)
        # the heads read the same pyramid and write disjoint buffers: independent branches of the plan
        for branch, (head, fn) in enumerate(((self.m.segheader, lambda: self.seg_head(feats[0], levels)),
                                             (self.m.detectheader, lambda: self.det_head(levels)),
                                             (self.m.laneheader, lambda: self.lane_head(levels))), start=1):
            if head is None:
                continue
            n0 = len(self.ops)
            fn()
            for op in self.ops[n0:]:
                op.branch = branch
        # ... and so are the two detection towers (branches must be contiguous and ascending: seg 1, reg 2, cls 3, lane 4); the
        # fused decode + NMS sits at the end of the classification tower's branch behind a wait for the regression tower
        for op in self.ops:
            b = getattr(op, "branch", 0)
            if b == 3 or (b == 2 and not op.name.startswith("det.reg")):
                op.branch = b + 1
        order = [getattr(op, "branch", 0) for op in self.ops]
        assert order == sorted(order), "head ops must be grouped by branch"
        return self


class Plan:
    """Compiled schedule for one input shape: static input, buffers, native plan handle."""

    def __init__(self, model, B, H, W, device, x=None, out_alloc=None):
        self.B, self.H, self.W, self.device = B, H, W, device
        self.x = x if x is not None else torch.zeros((B, 3, H, W), dtype=torch.float32, device=device)
        src = model.packing_source() if hasattr(model, "packing_source") else model
        self.builder = Builder(src, B, H, W, device, out_alloc=out_alloc).build(self.x)
        self.ops = self.builder.ops
        self.out = self.builder.out
        import ctypes
        self.handle = ctypes.c_void_p()
        nv.check(nv.lib.hn_plan_create(ctypes.byref(self.handle)))
        branches = bool(getattr(model, "head_branches", False))
        cur = 0
        for op in self.ops:
            if op.kind == "wait" and not branches:
                continue  # one stream: program order already is the dependency
            try:
                b = getattr(op, "branch", 0) if branches else 0
                if b != cur:
                    nv.check(nv.lib.hn_plan_set_branch(self.handle, b))
                    cur = b
                op.add_to(self.handle)
            except nv.NativeError as e:
                raise nv.NativeError("while adding op %s: %s" % (op.name, e))
        self.n_launches = nv.lib.hn_plan_num_launches(self.handle)
        self.graph_ready = False

    def run(self, stream_ptr):
        nv.check(nv.lib.hn_plan_run(self.handle, stream_ptr))

    def run_range(self, first, last, stream_ptr):
        nv.check(nv.lib.hn_plan_run_range(self.handle, first, last, stream_ptr))

    def capture(self, stream_ptr):
        nv.check(nv.lib.hn_plan_graph_capture(self.handle, stream_ptr))
        self.graph_ready = True

    def launch_graph(self, stream_ptr):
        nv.check(nv.lib.hn_plan_graph_launch(self.handle, stream_ptr))

    def __del__(self):
        try:
            if self.handle:
                nv.lib.hn_plan_destroy(self.handle)
        except Exception:
            pass


class SplitPlan:
    """The batch as two half-batch plans replayed on two streams inside ONE CUDA graph (fork / join).

    Most launches of the forward are latency-bound, not throughput-bound: a RegNet block is five small dependent
    kernels (1x1 GEMM, grouped 3x3, squeeze-excite gate, scale, 1x1 GEMM) that each occupy a fraction of the SMs for
    10-20 us.  Two independent half-batches interleave on the idle SMs and fill each other's tile-quantisation tails,
    while the wide, throughput-bound layers simply take turns.  Both halves write batch slices of the same output
    tensors, so callers see one [B, ...] result; the input is one static [B, 3, H, W] tensor as well.
    Measured on B200 at batch 32: 9.67 vs 9.40 ms/step -- the persistent conv CTAs take a whole SM's shared memory, so the
    halves' kernels never share an SM and every launch's fixed cost is paid twice.  Kept (off by default, HN_SPLIT=1)
    with its bit-identity test."""

    def __init__(self, model, B, H, W, device):
        assert B >= 2
        self.B, self.H, self.W, self.device = B, H, W, device
        self.x = torch.zeros((B, 3, H, W), dtype=torch.float32, device=device)
        self.out = {}
        b0 = (B + 1) // 2
        self.bounds = [(0, b0), (b0, B)]
        self.parts = []
        for lo, hi in self.bounds:
            def alloc(name, shape, dtype, lo=lo, hi=hi):
                if name not in self.out:
                    self.out[name] = torch.zeros((B,) + tuple(shape[1:]), dtype=dtype, device=device)
                return self.out[name][lo:hi]
            self.parts.append(Plan(model, hi - lo, H, W, device, x=self.x[lo:hi], out_alloc=alloc))
        self.ops = self.parts[0].ops  # (diagnostics: the op list of one half)
        self.n_launches = sum(p.n_launches for p in self.parts)
        self.side = torch.cuda.Stream(device)
        self.graph = None
        self.graph_ready = False

    def _run_forked(self, main):
        """main: torch stream.  Half 0 on `main`, half 1 on the side stream, joined back into `main`."""
        self.side.wait_stream(main)
        self.parts[0].run(main.cuda_stream)
        self.parts[1].run(self.side.cuda_stream)
        main.wait_stream(self.side)

    def run(self, stream_ptr):
        self._run_forked(torch.cuda.current_stream(self.device))

    def run_range(self, first, last, stream_ptr):  # diagnostics only: both halves, one after the other
        for p in self.parts:
            p.run_range(first, last, stream_ptr)

    def capture(self, stream_ptr):
        main = torch.cuda.current_stream(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            self._run_forked(torch.cuda.current_stream(self.device))
        self.graph = g
        self.graph_ready = True

    def launch_graph(self, stream_ptr):
        self.graph.replay()

