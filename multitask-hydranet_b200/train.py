"""Train-mode HydraNet: the reference's training forward / backward (model/train.py:241-269)
as ``torch.autograd.Function``s whose forward AND backward are native sm_100a kernels.

PyTorch supplies the tape (so ``loss.backward()``, ``torch.optim.*`` and DDP work unchanged on the facade -- the drop-in
contract of model.py), device memory and streams; every tensor operation of the network itself is a call into
``libhydranet_b200.so``:

  * convolutions, all flavours   forward and data gradient = ``hn_conv_fwd`` (tcgen05 implicit GEMM; dgrad is a convolution
                                 of the output gradient with the transposed / flipped filter), weight gradient =
                                 ``hn_conv_wgrad`` (tcgen05, K = pixels, MN-major operands straight from NHWC)
  * BatchNorm2d (batch stats)    ``hn_bn_train_fwd/bwd`` (anynet.py, common.py:95-99, detection.py:20-24, lanedetect.py:48)
  * depthwise 3x3                ``hn_dw_multi_fwd`` (forward and, with the flipped filter, data gradient), ``hn_dw_wgrad``
  * squeeze-excite               ``hn_col_reduce`` + ``hn_se_fc_fwd/bwd`` + ``hn_se_apply``            (anynet.py:39-47)
  * BiFPN fusion                 ``hn_wsum_swish_fwd``, ``hn_act_bwd``, ``hn_resample_fwd/bwd``        (bifpn.py:156-233)
  * seg decoder assembly         ``hn_seggather_fwd/bwd`` (ReflectionPad2d(1) of cat(up2(x), skip), segmentation.py:84-105)
  * head outputs                 fp32 tensors in the reference's layouts written by the GEMM epilogue; their gradients come
                                 back through ``hn_head_grad``
  * weight packing               ``hn_pack_weights``: fp32 parameters -> bf16 K-major matrices, ONE launch per step

Activations and activation gradients are bf16 NHWC; parameters, their gradients and all BatchNorm / squeeze-excite
statistics are fp32.  Parameter gradients are returned through autograd in the parameters' own layouts.
"""
import ctypes as C
import os

import torch
from torch.autograd import Function

from . import _native as nv
from .engine import choose_bn, choose_stages, choose_tile

BF = torch.bfloat16
_SCRATCH_FLOATS = 24 * 1024 * 1024  # 96 MB of partial-sum scratch per device


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _view(t):
    """hn_view of a [N,H,W,C] bf16 tensor (channel stride 1; other strides free: channel slices, phases)."""
    assert t.dim() == 4 and t.dtype == BF and t.stride(3) == 1, (tuple(t.shape), t.dtype, t.stride())
    return nv.View(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3], t.stride(0), t.stride(1), t.stride(2))


def _rows_view(t):
    """[1,1,R,C] view of a tensor whose pixels are equally spaced rows (2-D [R,C] or 4-D NHWC, channel slices allowed)."""
    m = _mat(t)
    return nv.View(m.ptr, 1, 1, m.rows, m.cols, 0, 0, m.ld)


def _mat(t):
    """hn_mat of a bf16 tensor [..., C] whose leading dims collapse to equally spaced rows."""
    assert t.dtype == BF and t.stride(-1) == 1, (t.dtype, t.stride())
    Cc = t.shape[-1]
    rows = t.numel() // Cc if Cc else 0
    ld = t.stride(-2) if t.dim() >= 2 else Cc
    exp = ld
    for d in range(t.dim() - 2, -1, -1):
        assert t.shape[d] == 1 or t.stride(d) == exp, ("rows are not equally spaced", tuple(t.shape), t.stride())
        exp *= t.shape[d]
    return nv.Mat(t.data_ptr(), rows, Cc, ld)


def _null_mat():
    return nv.Mat(None, 0, 0, 0)


def _null_view():
    return nv.View(None, 0, 0, 0, 0, 0, 0, 0)


def _phase(t, ry, rx):
    return t[:, ry::2, rx::2, :]


# --------------------------------------------------------------------------------------------------------------------
# packed weights
# --------------------------------------------------------------------------------------------------------------------
class Blocks:
    """One packed GEMM operand: a list of 64-wide K blocks (the conv kernel's taps) over a parameter tensor."""

    def __init__(self, rows, bn):
        self.rows, self.bn = rows, bn
        self.rows_pad = (rows + bn - 1) // bn * bn
        self.taps = []      # (src, dy, dx, c0)
        self.entries = []   # (src_elem_offset, s_r, s_c, cols, grouped)
        self.buf = None

    def add(self, src, dy, dx, c0, cols, off, s_r, s_c, grouped=0):
        self.taps.append((src, dy, dx, c0))
        self.entries.append((off, s_r, s_c, cols, grouped))


class ConvRec:
    """Static description of one convolution parameter: packed operands for forward and data gradient, tap geometry for the
    weight gradient."""

    def __init__(self, name, w, kind, stride=1, src_channels=None):
        self.name, self.w, self.kind, self.stride = name, w, kind, stride
        cout, cin_g, kh, kw = w.shape
        self.cout = cout
        self.grouped = 1 if kind == "g3" else 0
        self.cin = cout if self.grouped else cin_g
        self.src_channels = list(src_channels) if src_channels else [self.cin]
        assert sum(self.src_channels) == self.cin
        self.s_co, self.s_ci = cin_g * kh * kw, kh * kw
        self.fwd = None      # Blocks (rows = cout)
        self.dgrad = None    # Blocks (rows = cin) or, for the strided grouped conv, {(ry, rx): Blocks}
        self.wg_taps = []    # (src, dy, dx, c0, tap_off, tap_cin)
        getattr(self, "_build_" + kind)()

    # -- 1x1, stride 1 (several concatenated sources allowed) / stride 2 (one source, read through its (0,0) phase) --
    def _build_pw(self):
        bn = choose_bn(self.cout)
        f = Blocks(self.cout, bn)
        woff = 0
        for s, cs in enumerate(self.src_channels):
            for c0 in range(0, cs, 64):
                cols = min(64, cs - c0)
                f.add(s, 0, 0, c0, cols, (woff + c0) * self.s_ci, self.s_co, self.s_ci)
                self.wg_taps.append((s, 0, 0, c0, (woff + c0) * self.s_ci, cols))
            woff += cs
        self.fwd = f
        d = Blocks(self.cin, choose_bn(self.cin))
        for co0 in range(0, self.cout, 64):
            d.add(0, 0, 0, co0, min(64, self.cout - co0), co0 * self.s_co, self.s_ci, self.s_co)
        self.dgrad = d

    _build_pw_s2 = _build_pw

    # -- dense 3x3 over a pre-padded input (taps 0..2), segmentation.py:32-48 --
    def _build_c3(self):
        f = Blocks(self.cout, choose_bn(self.cout))
        for ky in range(3):
            for kx in range(3):
                for c0 in range(0, self.cin, 64):
                    cols = min(64, self.cin - c0)
                    f.add(0, ky, kx, c0, cols, c0 * self.s_ci + ky * 3 + kx, self.s_co, self.s_ci)
                    self.wg_taps.append((0, ky, kx, c0, c0 * self.s_ci + ky * 3 + kx, cols))
        self.fwd = f
        d = Blocks(self.cin, choose_bn(self.cin))  # gradient of the PADDED input: dpad[Y, X] = sum dy[Y - ky, X - kx] . W[ky, kx]
        for ky in range(3):
            for kx in range(3):
                for co0 in range(0, self.cout, 64):
                    d.add(0, -ky, -kx, co0, min(64, self.cout - co0), co0 * self.s_co + ky * 3 + kx, self.s_ci, self.s_co)
        self.dgrad = d

    # -- grouped 3x3, group width 8, zero pad 1, stride 1 or 2 (anynet.py:33-37) --
    def _build_g3(self):
        Cc = self.cout
        assert self.w.shape[1] == 8 and Cc % 8 == 0, "grouped conv path assumes group width 8"
        ph = {0: (1, -1), 1: (0, 0), 2: (1, 0)}  # stride 2: input row 2y+ky-1 = 2(y+a)+r -> ky: (r, a)
        f = Blocks(Cc, 64)
        for ky in range(3):
            for kx in range(3):
                if self.stride == 1:
                    s, dy, dx = 0, ky - 1, kx - 1
                else:
                    s, dy, dx = ph[ky][0] * 2 + ph[kx][0], ph[ky][1], ph[kx][1]
                f.add(s, dy, dx, 0, 64, ky * 3 + kx, self.s_co, self.s_ci, grouped=1)
                for half in (0, 64):
                    if half < Cc:
                        self.wg_taps.append((s, dy, dx, half, ky * 3 + kx, 64))
        self.fwd = f
        if self.stride == 1:
            d = Blocks(Cc, 64)
            for ky in range(3):
                for kx in range(3):
                    d.add(0, 1 - ky, 1 - kx, 0, 64, ky * 3 + kx, self.s_co, self.s_ci, grouped=2)
            self.dgrad = d
        else:
            # dx[2u+ry, 2v+rx]: ry = 0 <- ky = 1 (y = u); ry = 1 <- ky = 0 (y = u + 1), ky = 2 (y = u)
            ks = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}
            self.dgrad = {}
            for ry in range(2):
                for rx in range(2):
                    d = Blocks(Cc, 64)
                    for ky, sy in ks[ry]:
                        for kx, sx in ks[rx]:
                            d.add(0, sy, sx, 0, 64, ky * 3 + kx, self.s_co, self.s_ci, grouped=2)
                    self.dgrad[(ry, rx)] = d

    def all_blocks(self):
        out = [self.fwd]
        out += list(self.dgrad.values()) if isinstance(self.dgrad, dict) else [self.dgrad]
        return out


class TrainState:
    """Per (model, device): conv records, packed-weight buffers, the device-side pack table, scratch."""

    def __init__(self, model, device):
        self.device = device
        self.recs = {}
        self.scratch = torch.empty(_SCRATCH_FLOATS, dtype=torch.float32, device=device)
        self.bn_seen = []
        self.side_stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self.side_wgrad = self.side_stream is not None and os.environ.get("HN_SIDE_WGRAD", "1") != "0"
        self.side_dirty = False
        if model is not None:
            self._build(model)
            self._upload_table()

    def rec(self, name, w, kind, **kw):
        r = self.recs.get(name)
        if r is None:
            assert w.dtype == torch.float32 and w.is_contiguous(), name
            r = self.recs[name] = ConvRec(name, w, kind, **kw)
        return r

    def _build(self, m):
        net = m.backbone.net
        cin = 32
        for s in range(m.backbone.stage_num):
            for bi, blk in enumerate(getattr(net, "stage_%d" % s).blocks.children()):
                p = "backbone.s%d.b%d" % (s, bi)
                self.rec(p + ".c1", blk.conv_block_1[0].weight, "pw")
                self.rec(p + ".c2", blk.conv_block_2[0].weight, "g3", stride=blk.stride)
                self.rec(p + ".c3", blk.conv_block_3[0].weight, "pw")
                if blk.shortcut is not None:
                    self.rec(p + ".sc", blk.shortcut[0].weight, "pw_s2" if blk.stride == 2 else "pw")
        for ci, cell in enumerate(m.neck.bifpn.children()):
            p = "neck.c%d" % ci
            for n in cell.NODES:
                self.rec(p + "." + n, getattr(cell, n).pointwise_conv.conv.weight, "pw")
            if cell.first_time:
                for n in ("p3_down_channel", "p4_down_channel", "p5_down_channel", "p4_down_channel_2", "p5_down_channel_2",
                          "p6_down_channel", "p5_to_p6"):
                    if hasattr(cell, n):
                        self.rec(p + "." + n, getattr(cell, n)[0].conv.weight, "pw")
        if m.segheader is not None:
            for i, d in enumerate(m.segheader.decoder.children()):
                conv = d.conv.conv if hasattr(d.conv, "conv") else d.conv
                self.rec("seg.d%d" % i, conv.weight, "c3")
        if m.detectheader is not None:
            for tn, tower in (("reg", m.detectheader.regressor), ("cls", m.detectheader.classifier)):
                for i in range(tower.num_layers):
                    self.rec("det.%s.%d" % (tn, i), tower.conv_list[i].pointwise_conv.conv.weight, "pw")
                self.rec("det.%s.hdr" % tn, tower.header.pointwise_conv.conv.weight, "pw")
        if m.laneheader is not None:
            lh = m.laneheader
            Cn = m.fpn_num_filters
            for bn_, br in (("cls", lh.conv_cls_conv), ("up", lh.conv_up_conv), ("down", lh.conv_down_conv)):
                self.rec("lane.%s.hid" % bn_, br[0].weight, "pw", src_channels=[Cn] * 4)
                self.rec("lane.%s.out" % bn_, br[3].weight, "pw")

    def _upload_table(self):
        entries, max_rows = [], 1
        for r in self.recs.values():
            for b in r.all_blocks():
                b.buf = torch.zeros((b.rows_pad, 64 * len(b.entries)), dtype=BF, device=self.device)
                max_rows = max(max_rows, b.rows_pad)
                for k, (off, s_r, s_c, cols, grouped) in enumerate(b.entries):
                    entries.append(nv.PackEntry(r.w.data_ptr() + 4 * off, b.buf.data_ptr() + 2 * 64 * k, b.buf.shape[1], b.rows, b.rows_pad,
                                                cols, s_r, s_c, grouped, 0))
        arr = (nv.PackEntry * len(entries))(*entries)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = raw.to(self.device)
        self.n_entries, self.max_rows = len(entries), max_rows
        self.sig = tuple(r.w.data_ptr() for r in self.recs.values())
        self.grad_offset, off = {}, 0
        for r in self.recs.values():
            self.grad_offset[r.name] = off
            off += (r.w.numel() + 63) // 64 * 64  # 256-byte aligned slices
        self.grad_elems = off
        self.token = StepToken(self)

    def pack(self):
        """fp32 parameters -> bf16 GEMM operands (forward and data-gradient matrices of every conv): one launch."""
        nv.check(nv.lib.hn_pack_weights(self.table.data_ptr(), self.n_entries, self.max_rows, _stream(self.device)))

    def valid_for(self, model):
        return self.sig == tuple(r.w.data_ptr() for r in self.recs.values())


class StepToken:
    """Per-forward handle: the backward of THIS forward allocates one zeroed fp32 arena for all convolution weight gradients
    (the split-K kernel accumulates into zeros) instead of ~150 separately filled tensors; the gradients handed to autograd are
    views of it (a fresh arena per step: nothing aliases a gradient the caller may still hold from the previous step)."""

    def __init__(self, st):
        self.st, self.arena, self.taken = st, None, set()

    def dw(self, rec):
        if rec.name in self.taken:  # a second backward through the same forward (retain_graph) or a re-used state: own buffer
            return torch.zeros_like(rec.w)
        self.taken.add(rec.name)
        if self.arena is None:
            self.arena = torch.zeros(self.st.grad_elems, dtype=torch.float32, device=self.st.device)
        off = self.st.grad_offset[rec.name]
        return self.arena[off:off + rec.w.numel()].view(rec.w.shape)


def get_state(model, device):
    st = getattr(model, "_train_state", None)
    if st is None or st.device != device or not st.valid_for(model):
        st = TrainState(model, device)
        object.__setattr__(model, "_train_state", st)
    return st


# --------------------------------------------------------------------------------------------------------------------
# GEMM launch helpers
# --------------------------------------------------------------------------------------------------------------------
def _conv_desc(srcs, blocks, *, flat, cout, out_ptr, out_strides, n_img=0, out_h=0, out_w=0, flat_hw=0, tile=(8, 16), bias=None, act=nv.ACT_NONE,
               out_fp32=0, out_scale=1, out_oy=0, out_ox=0, grouped=0, groups=None):
    d = nv.ConvDesc()
    for i, v in enumerate(srcs):
        d.src[i] = v
    d.n_src = len(srcs)
    d.weight = blocks.buf.data_ptr()
    d.w_rows = blocks.rows_pad
    d.num_taps = len(blocks.taps)
    assert d.num_taps <= nv.HN_MAX_TAPS
    for i, (s, dy, dx, c0) in enumerate(blocks.taps):
        d.taps[i] = nv.Tap(s, dy, dx, 0, c0, 0)
    d.flat = 1 if flat else 0
    d.tile_h, d.tile_w = tile
    d.n_img, d.out_h, d.out_w, d.flat_hw = n_img, out_h, out_w, flat_hw
    d.cout, d.bn = cout, blocks.bn
    d.stages = choose_stages(blocks.bn, len(blocks.taps))
    d.bias = bias
    d.act, d.epi = act, nv.EPI_STD
    d.out = out_ptr
    d.out_fp32 = out_fp32
    d.out_stride_n, d.out_stride_y, d.out_stride_x = out_strides
    d.out_scale, d.out_oy, d.out_ox, d.halo = out_scale, out_oy, out_ox, nv.HALO_NONE
    d.grouped = grouped
    if groups is not None:
        ends, hws, bases = groups
        d.n_groups = len(ends)
        d.group_addr = 1
        for i in range(len(ends)):
            d.group_end[i], d.group_hw[i], d.group_out_base[i] = ends[i], hws[i], bases[i]
    return d


def _gemm_rows(dev, x, blocks, n_out, out, bias=None, act=nv.ACT_NONE):
    """out[R, n_out] (bf16 rows) = act(x[R, K] . W^T + bias): 1x1 convolution / its data gradient in flat mode."""
    m = _mat(out)
    d = _conv_desc([_rows_view(x)], blocks, flat=True, cout=n_out, out_ptr=m.ptr, out_strides=(m.rows * m.ld, 0, m.ld), flat_hw=max(m.rows, 1),
                   bias=bias.data_ptr() if bias is not None else None, act=act)
    nv.check(nv.lib.hn_conv_fwd(C.byref(d), _stream(dev)))


def _gemm_spatial(dev, srcs, blocks, n_out, out, tile_hw, bias=None, act=nv.ACT_NONE, grouped=0, out_scale=1, oy=0, ox=0, out_fp32=0):
    """Spatially tiled conv: `out` is the full output tensor [N,H,W,C]; the tile space is `tile_hw` (= output size / out_scale)."""
    th, tw = tile_hw
    es = 1
    d = _conv_desc(srcs, blocks, flat=False, cout=n_out, out_ptr=out.data_ptr(), out_strides=(out.stride(0) * es, out.stride(1) * es, out.stride(2) * es),
                   n_img=out.shape[0], out_h=th, out_w=tw, tile=choose_tile(th, tw), bias=bias.data_ptr() if bias is not None else None, act=act,
                   grouped=grouped, out_scale=out_scale, out_oy=oy, out_ox=ox, out_fp32=out_fp32)
    nv.check(nv.lib.hn_conv_fwd(C.byref(d), _stream(dev)))


_PROFILE_NAMES = bool(os.environ.get("HN_PROFILE_NAMES"))


def _wgrad(st, rec, dy_t, src_ts, dy_view, src_views, flat, tile, dw):
    """Weight gradient of one convolution.  Nothing downstream in backward depends on it (only the optimizer does), and most of
    these launches are small and latency-bound, so they go to a SIDE stream: they fill the SMs the main chain (data gradient ->
    BatchNorm backward -> next layer) leaves idle.  Joined back in StemConv.backward (the last node of every backward) and in
    TrainStep.  Not used when the parameter already holds a gradient (autograd would add to it on the main stream at once)."""
    dev = dw.device
    side = st.side_stream if (st.side_wgrad and rec.w.grad is None) else None
    if side is None:
        return _wgrad_named(dev, rec, dy_view, src_views, flat, tile, dw, _stream(dev))
    main = torch.cuda.current_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        _wgrad_named(dev, rec, dy_view, src_views, flat, tile, dw, side.cuda_stream)
    for t in (dy_t, dw) + tuple(src_ts):
        t.record_stream(side)
    if not st.side_dirty:
        st.side_dirty = True
        # join when this backward pass is over, whatever its last node is (runs on the thread that called backward())
        torch.autograd.Variable._execution_engine.queue_callback(lambda: join_side_stream(st, dev))


def join_side_stream(st, dev):
    if st.side_dirty:
        torch.cuda.current_stream(dev).wait_stream(st.side_stream)
        st.side_dirty = False


_NVTX = bool(os.environ.get("HN_NVTX"))


def _wgrad_named(dev, rec, dy_view, src_views, flat, tile, dw, stream):
    if _NVTX:  # tools/profile_train_step.py under `ncu --nvtx --nvtx-include "wgrad:<layer>/"`
        torch.cuda.nvtx.range_push("wgrad:" + rec.name)
        try:
            return _wgrad_launch(dev, rec, dy_view, src_views, flat, tile, dw, stream)
        finally:
            torch.cuda.nvtx.range_pop()
    if _PROFILE_NAMES:  # tools/profile_train.py: per-layer device time of the weight-gradient kernel
        with torch.profiler.record_function("wgrad:" + rec.name):
            return _wgrad_launch(dev, rec, dy_view, src_views, flat, tile, dw, stream)
    return _wgrad_launch(dev, rec, dy_view, src_views, flat, tile, dw, stream)


def _wgrad_launch(dev, rec, dy_view, src_views, flat, tile, dw, stream):
    d = nv.WgradDesc()
    d.dy = dy_view
    for i, v in enumerate(src_views):
        d.src[i] = v
    d.n_src = len(src_views)
    d.num_taps = len(rec.wg_taps)
    assert d.num_taps <= nv.HN_MAX_TAPS, rec.name
    for i, (s, dy, dx, c0, off, cin) in enumerate(rec.wg_taps):
        d.taps[i] = nv.Tap(s, dy, dx, 0, c0, 0)
        d.tap_off[i] = off
        d.tap_cin[i] = cin
    d.flat = 1 if flat else 0
    d.tile_h, d.tile_w = tile
    d.cout = rec.cout
    d.s_co, d.s_ci = rec.s_co, rec.s_ci
    d.grouped = rec.grouped
    d.dw = dw.data_ptr()
    nv.check(nv.lib.hn_conv_wgrad(C.byref(d), stream))


def _colsum(st, t, valid=None):
    """fp32 [C] column sums of a bf16 row tensor (bias gradients)."""
    out = torch.empty(t.shape[-1], dtype=torch.float32, device=t.device)
    m = _mat(t)
    nv.check(nv.lib.hn_col_reduce(C.byref(m), None, 0, 0, out.data_ptr(), None, 1.0, st.scratch.data_ptr(), st.scratch.numel() * 4, _stream(t.device)))
    return out if valid is None else out[:valid]


def _act_bwd(dy, ref, act, scaled=(), w=None):
    dz = torch.empty_like(dy, memory_format=torch.contiguous_format)
    d = nv.ActBwdDesc()
    d.dy, d.dz, d.act = _mat(dy), _mat(dz), act
    d.ref = _mat(ref) if act != nv.ACT_NONE else _null_mat()
    d.n_scaled = len(scaled)
    for k, t in enumerate(scaled):
        d.scaled[k] = _mat(t)
    d.w = w.data_ptr() if w is not None else None
    nv.check(nv.lib.hn_act_bwd(C.byref(d), _stream(dy.device)))
    return dz


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------------------------------
# autograd functions
# --------------------------------------------------------------------------------------------------------------------
class StemConv(Function):
    """anynet.py:11: 3x3 s2 p1 conv 3 -> 32 on the fp32 NCHW input (raw output; BatchNorm follows)."""

    @staticmethod
    def forward(ctx, st, x, w):
        N, _, H, W = x.shape
        out = torch.empty((N, (H + 1) // 2, (W + 1) // 2, 32), dtype=BF, device=x.device)
        wk = w.detach().permute(1, 2, 3, 0).reshape(27, 32).contiguous()
        zero = torch.zeros(32, dtype=torch.float32, device=x.device)
        d = nv.StemDesc(x.data_ptr(), N, H, W, wk.data_ptr(), zero.data_ptr(), _view(out), 1)
        nv.check(nv.lib.hn_stem_fwd(C.byref(d), _stream(x.device)))
        ctx.st = st
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dz):
        (x,) = ctx.saved_tensors
        st = ctx.st
        dz = _c(dz)
        N, _, H, W = x.shape
        dw = torch.empty((32, 3, 3, 3), dtype=torch.float32, device=x.device)
        v = _view(dz)
        nv.check(nv.lib.hn_stem_wgrad(x.data_ptr(), N, H, W, C.byref(v), dw.data_ptr(), st.scratch.data_ptr(), st.scratch.numel() * 4, _stream(x.device)))
        join_side_stream(st, x.device)  # every other backward node has run by now: the weight gradients on the side stream are complete
        return None, None, dw


class Conv1x1(Function):
    """1x1 convolution over one or several channel-concatenated sources (the concat is never materialised), optional bias.
    Inputs / output: bf16 [..., C] with equally spaced rows."""

    @staticmethod
    def forward(ctx, st, rec, bias, w, *xs):
        dev = xs[0].device
        lead = xs[0].shape[:-1]
        out = torch.empty(tuple(lead) + (rec.cout,), dtype=BF, device=dev)
        m = _mat(out)
        d = _conv_desc([_rows_view(x) for x in xs], rec.fwd, flat=True, cout=rec.cout, out_ptr=m.ptr, out_strides=(m.rows * m.ld, 0, m.ld),
                       flat_hw=max(m.rows, 1), bias=bias.detach().data_ptr() if bias is not None else None)
        nv.check(nv.lib.hn_conv_fwd(C.byref(d), _stream(dev)))
        ctx.st, ctx.rec, ctx.has_bias, ctx.tok = st, rec, bias is not None, st.token
        ctx.save_for_backward(*xs)
        return out

    @staticmethod
    def backward(ctx, dy):
        xs = ctx.saved_tensors
        st, rec = ctx.st, ctx.rec
        dev = dy.device
        dy = _c(dy)
        dx_full = torch.empty(tuple(dy.shape[:-1]) + (rec.cin,), dtype=BF, device=dev)
        _gemm_rows(dev, dy, rec.dgrad, rec.cin, dx_full)
        dw = ctx.tok.dw(rec)
        _wgrad(st, rec, dy, xs, _rows_view(dy), [_rows_view(x) for x in xs], True, (1, 128), dw)
        db = _colsum(st, dy) if ctx.has_bias else None
        if len(xs) == 1:
            dxs = (dx_full,)
        else:
            dxs, c0 = [], 0
            for cs in rec.src_channels:
                dxs.append(dx_full[..., c0:c0 + cs])
                c0 += cs
        return (None, None, db, dw) + tuple(dxs)


class Conv1x1S2(Function):
    """Stride-2 1x1 shortcut convolution (anynet.py:54-58): reads the (0,0) phase of its input."""

    @staticmethod
    def forward(ctx, st, rec, w, x):
        N, H, W, _ = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = torch.empty((N, Ho, Wo, rec.cout), dtype=BF, device=x.device)
        _gemm_spatial(x.device, [_view(_phase(x, 0, 0))], rec.fwd, rec.cout, out, (Ho, Wo))
        ctx.st, ctx.rec, ctx.tok = st, rec, st.token
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        rec = ctx.rec
        dy = _c(dy)
        dx = torch.zeros_like(x)  # only the (0,0) phase receives gradient
        _gemm_spatial(x.device, [_view(dy)], rec.dgrad, rec.cin, dx, (dy.shape[1], dy.shape[2]), out_scale=2)
        dw = ctx.tok.dw(rec)
        _wgrad(ctx.st, rec, dy, [x], _view(dy), [_view(_phase(x, 0, 0))], False, choose_tile(dy.shape[1], dy.shape[2]), dw)
        return None, None, dw, dx


class GroupedConv3x3(Function):
    """anynet.py:33-37: 3x3, groups of 8 channels, zero pad 1, stride 1 or 2, as block-diagonal tensor-core tiles."""

    @staticmethod
    def _srcs(x, stride):
        return [_view(x)] if stride == 1 else [_view(_phase(x, 0, 0)), _view(_phase(x, 0, 1)), _view(_phase(x, 1, 0)), _view(_phase(x, 1, 1))]

    @staticmethod
    def forward(ctx, st, rec, w, x):
        N, H, W, Cc = x.shape
        s = rec.stride
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        assert s == 1 or (H % 2 == 0 and W % 2 == 0), "strided grouped conv expects even maps"
        out = torch.empty((N, Ho, Wo, Cc), dtype=BF, device=x.device)
        _gemm_spatial(x.device, GroupedConv3x3._srcs(x, s), rec.fwd, Cc, out, (Ho, Wo), grouped=1)
        ctx.st, ctx.rec, ctx.tok = st, rec, st.token
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        rec = ctx.rec
        dev = x.device
        dy = _c(dy)
        dx = torch.empty_like(x)
        if rec.stride == 1:
            _gemm_spatial(dev, [_view(dy)], rec.dgrad, rec.cin, dx, (x.shape[1], x.shape[2]), grouped=1)
        else:
            for (ry, rx), blocks in rec.dgrad.items():
                _gemm_spatial(dev, [_view(dy)], blocks, rec.cin, dx, (x.shape[1] // 2, x.shape[2] // 2), grouped=1, out_scale=2, oy=ry, ox=rx)
        dw = ctx.tok.dw(rec)
        _wgrad(ctx.st, rec, dy, [x], _view(dy), GroupedConv3x3._srcs(x, rec.stride), False, choose_tile(dy.shape[1], dy.shape[2]), dw)
        return None, None, dw, dx


class Conv3x3Padded(Function):
    """Dense 3x3 over an already padded input [N,H+2,W+2,C] (+ bias, + ELU): the segmentation decoder's ConvBlock / Conv3x3
    (segmentation.py:16-48).  ``logits``: fp32 NHWC output without activation (decoder.8)."""

    @staticmethod
    def forward(ctx, st, rec, act, logits, w, bias, xp):
        N, Hp, Wp, _ = xp.shape
        H, W = Hp - 2, Wp - 2
        dev = xp.device
        out = torch.empty((N, H, W, rec.cout), dtype=torch.float32 if logits else BF, device=dev)
        _gemm_spatial(dev, [_view(xp)], rec.fwd, rec.cout, out, (H, W), bias=bias.detach(), act=act, out_fp32=1 if logits else 0)
        ctx.st, ctx.rec, ctx.act, ctx.logits, ctx.tok = st, rec, act, logits, st.token
        ctx.save_for_backward(xp, out)
        return out

    @staticmethod
    def backward(ctx, dy):
        xp, out = ctx.saved_tensors
        st, rec = ctx.st, ctx.rec
        dev = xp.device
        N, Hp, Wp, _ = xp.shape
        if ctx.logits:  # fp32 NHWC gradient of the logits -> bf16 rows padded to 8 columns
            dy = _c(dy)
            cp = (rec.cout + 7) // 8 * 8
            dz = torch.empty((N, Hp - 2, Wp - 2, cp), dtype=BF, device=dev)
            hd = nv.HeadGradDesc()
            hd.dout, hd.out, hd.act, hd.cols_valid = dy.data_ptr(), None, nv.ACT_NONE, rec.cout
            hd.stride_n, hd.stride_pix, hd.stride_c, hd.rows_per_img = (Hp - 2) * (Wp - 2) * rec.cout, rec.cout, 1, (Hp - 2) * (Wp - 2)
            hd.n_groups, hd.dz = 0, _mat(dz)
            nv.check(nv.lib.hn_head_grad(C.byref(hd), _stream(dev)))
            db = dy.sum(dim=(0, 1, 2))
        else:
            dz = _act_bwd(_c(dy), out, ctx.act)
            db = _colsum(st, dz)
        dxp = torch.empty_like(xp)
        _gemm_spatial(dev, [_view(dz)], rec.dgrad, rec.cin, dxp, (Hp, Wp))
        dw = ctx.tok.dw(rec)
        _wgrad(st, rec, dz, [xp], _view(dz), [_view(xp)], False, choose_tile(Hp - 2, Wp - 2), dw)
        return None, None, None, None, dw, db, dxp


class SegGather(Function):
    """ReflectionPad2d(1)(cat(up2(low), skip)) -- either part optional (segmentation.py:84-105)."""

    @staticmethod
    def forward(ctx, low, skip):
        ref = skip if skip is not None else low
        N = ref.shape[0]
        H, W = (skip.shape[1], skip.shape[2]) if skip is not None else (low.shape[1] * 2, low.shape[2] * 2)
        Cc = (low.shape[3] if low is not None else 0) + (skip.shape[3] if skip is not None else 0)
        out = torch.empty((N, H + 2, W + 2, Cc), dtype=BF, device=ref.device)
        d = nv.SegGatherDesc(_view(low) if low is not None else _null_view(), _view(skip) if skip is not None else _null_view(), _view(out),
                             _null_view(), _null_view())
        nv.check(nv.lib.hn_seggather_fwd(C.byref(d), _stream(ref.device)))
        ctx.shapes = (None if low is None else tuple(low.shape), None if skip is None else tuple(skip.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        ls, ss = ctx.shapes
        dlow = torch.empty(ls, dtype=BF, device=dout.device) if ls else None
        dskip = torch.empty(ss, dtype=BF, device=dout.device) if ss else None
        d = nv.SegGatherDesc(_null_view(), _null_view(), _view(dout), _view(dlow) if ls else _null_view(), _view(dskip) if ss else _null_view())
        nv.check(nv.lib.hn_seggather_bwd(C.byref(d), _stream(dout.device)))
        return dlow, dskip


class BatchNormAct(Function):
    """nn.BatchNorm2d in training mode (batch statistics, running-stat update) + optional residual add + activation.
    ``seg_end``: row segments with their own BatchNorm (the pyramid levels of a detection tower)."""

    @staticmethod
    def forward(ctx, st, bns, act, seg_end, z, res, *gb):
        n = len(bns)
        gammas, betas = gb[:n], gb[n:]
        dev = z.device
        Cc = z.shape[-1]
        y = torch.empty_like(z)
        stats = torch.empty((n, 4, Cc), dtype=torch.float32, device=dev)
        d = nv.BnDesc()
        d.z, d.y, d.n_seg = _mat(z), _mat(y), n
        rows = d.z.rows
        ends = list(seg_end) if seg_end is not None else [rows]
        for i in range(n):
            d.seg_end[i] = ends[i]
            d.gamma[i], d.beta[i] = gammas[i].detach().data_ptr(), betas[i].detach().data_ptr()
            d.running_mean[i], d.running_var[i] = bns[i].running_mean.data_ptr(), bns[i].running_var.data_ptr()
        d.eps, d.momentum = bns[0].eps, bns[0].momentum
        d.stats, d.act = stats.data_ptr(), act
        d.res = _mat(res) if res is not None else _null_mat()
        d.scratch, d.scratch_bytes = st.scratch.data_ptr(), st.scratch.numel() * 4
        nv.check(nv.lib.hn_bn_train_fwd(C.byref(d), _stream(dev)))
        st.bn_seen.extend(bns)
        ctx.st, ctx.n, ctx.act, ctx.ends, ctx.has_res = st, n, act, ends, res is not None
        ctx.save_for_backward(z, y if act == nv.ACT_RELU else z, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, y, stats = ctx.saved_tensors
        st, n = ctx.st, ctx.n
        dev = z.device
        dy = _c(dy)
        Cc = z.shape[-1]
        dz = torch.empty_like(z)
        dres = torch.empty_like(z) if ctx.has_res else None
        dg = torch.empty((n, Cc), dtype=torch.float32, device=dev)
        db = torch.empty((n, Cc), dtype=torch.float32, device=dev)
        d = nv.BnDesc()
        d.z, d.y, d.dy, d.dz, d.n_seg = _mat(z), _mat(y), _mat(dy), _mat(dz), n
        d.dres = _mat(dres) if dres is not None else _null_mat()
        for i in range(n):
            d.seg_end[i] = ctx.ends[i]
            d.dgamma[i], d.dbeta[i] = dg[i].data_ptr(), db[i].data_ptr()
        d.stats, d.act = stats.data_ptr(), ctx.act
        d.scratch, d.scratch_bytes = st.scratch.data_ptr(), st.scratch.numel() * 4
        nv.check(nv.lib.hn_bn_train_bwd(C.byref(d), _stream(dev)))
        return (None, None, None, None, dz, dres) + tuple(dg[i] for i in range(n)) + tuple(db[i] for i in range(n))


class SqueezeExcite(Function):
    """anynet.py:39-47,68-69: y = x * sigmoid(W2 relu(W1 mean(x) + b1) + b2)."""

    @staticmethod
    def forward(ctx, st, x, w1, b1, w2, b2):
        N, H, W, Cc = x.shape
        S = w1.shape[0]
        dev = x.device
        mean = torch.empty((N, Cc), dtype=torch.float32, device=dev)
        h = torch.empty((N, S), dtype=torch.float32, device=dev)
        gate = torch.empty((N, Cc), dtype=torch.float32, device=dev)
        xm = _mat(x)
        nv.check(nv.lib.hn_col_reduce(C.byref(xm), None, 0, H * W, mean.data_ptr(), None, 1.0 / (H * W), st.scratch.data_ptr(), st.scratch.numel() * 4, _stream(dev)))
        d = nv.SeFcDesc()
        d.N, d.C, d.S, d.mean = N, Cc, S, mean.data_ptr()
        d.w1, d.b1, d.w2, d.b2 = w1.detach().data_ptr(), b1.detach().data_ptr(), w2.detach().data_ptr(), b2.detach().data_ptr()
        d.h, d.gate = h.data_ptr(), gate.data_ptr()
        nv.check(nv.lib.hn_se_fc_fwd(C.byref(d), _stream(dev)))
        y = torch.empty_like(x)
        ym = _mat(y)
        nv.check(nv.lib.hn_se_apply(C.byref(xm), gate.data_ptr(), None, H * W, C.byref(ym), _stream(dev)))
        ctx.st = st
        ctx.save_for_backward(x, mean, h, gate, w1, w2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, h, gate, w1, w2 = ctx.saved_tensors
        st = ctx.st
        N, H, W, Cc = x.shape
        S = w1.shape[0]
        dev = x.device
        dy = _c(dy)
        dgate = torch.empty((N, Cc), dtype=torch.float32, device=dev)
        dym, xm = _mat(dy), _mat(x)
        nv.check(nv.lib.hn_col_reduce(C.byref(dym), C.byref(xm), 1, H * W, dgate.data_ptr(), None, 1.0, st.scratch.data_ptr(), st.scratch.numel() * 4, _stream(dev)))
        dmean = torch.empty((N, Cc), dtype=torch.float32, device=dev)
        dw1, db1 = torch.empty_like(w1), torch.empty(S, dtype=torch.float32, device=dev)
        dw2, db2 = torch.empty_like(w2), torch.empty(Cc, dtype=torch.float32, device=dev)
        tmp = torch.empty((N, Cc + S), dtype=torch.float32, device=dev)
        d = nv.SeFcDesc()
        d.N, d.C, d.S, d.mean = N, Cc, S, mean.data_ptr()
        d.w1, d.w2, d.h, d.gate = w1.data_ptr(), w2.data_ptr(), h.data_ptr(), gate.data_ptr()
        d.dgate, d.dmean, d.tmp = dgate.data_ptr(), dmean.data_ptr(), tmp.data_ptr()
        d.dw1, d.db1, d.dw2, d.db2 = dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr()
        nv.check(nv.lib.hn_se_fc_bwd(C.byref(d), _stream(dev)))
        dmean.mul_(1.0 / (H * W))
        dx = torch.empty_like(x)
        dxm = _mat(dx)
        nv.check(nv.lib.hn_se_apply(C.byref(dym), gate.data_ptr(), dmean.data_ptr(), H * W, C.byref(dxm), _stream(dev)))
        return None, dx, dw1, db1, dw2, db2


class Depthwise3x3(Function):
    """Depthwise 3x3, zero pad 1, no bias (common.py:91-93), over one or several maps that share the filter; several maps
    produce one stacked [rows, C] output (the detection towers run all pyramid levels as one matrix)."""

    @staticmethod
    def _run(xs, outs, w9):
        d = nv.DwMultiDesc()
        d.n = len(xs)
        for i, (a, b) in enumerate(zip(xs, outs)):
            d.in_[i], d.out[i] = _view(a), _view(b)
        d.dw = w9.data_ptr()
        nv.check(nv.lib.hn_dw_multi_fwd(C.byref(d), _stream(xs[0].device)))

    @staticmethod
    def _level_views(stacked, shapes):
        views, r0 = [], 0
        for (N, H, W, Cc) in shapes:
            views.append(stacked[r0:r0 + N * H * W].view(N, H, W, Cc))
            r0 += N * H * W
        return views

    @staticmethod
    def forward(ctx, st, stack, w, *xs):
        Cc = xs[0].shape[3]
        shapes = [tuple(x.shape) for x in xs]
        w9 = w.detach().reshape(Cc, 9).t().contiguous()
        if stack:
            out = torch.empty((sum(s[0] * s[1] * s[2] for s in shapes), Cc), dtype=BF, device=xs[0].device)
            outs = Depthwise3x3._level_views(out, shapes)
        else:
            out = torch.empty_like(xs[0], memory_format=torch.contiguous_format)
            outs = [out]
        Depthwise3x3._run(xs, outs, w9)
        ctx.st, ctx.stack, ctx.shapes = st, stack, shapes
        ctx.save_for_backward(w, *xs)
        return out

    @staticmethod
    def backward(ctx, dy):
        w, *xs = ctx.saved_tensors
        st = ctx.st
        dev = dy.device
        dy = _c(dy)
        Cc = xs[0].shape[3]
        dys = Depthwise3x3._level_views(dy, ctx.shapes) if ctx.stack else [dy]
        dxs = [torch.empty(s, dtype=BF, device=dev) for s in ctx.shapes]
        Depthwise3x3._run(dys, dxs, w.detach().reshape(Cc, 9).flip(1).t().contiguous())
        dw9 = torch.empty((9, Cc), dtype=torch.float32, device=dev)
        for i, (x, g) in enumerate(zip(xs, dys)):
            xv, gv = _view(x), _view(g)
            nv.check(nv.lib.hn_dw_wgrad(C.byref(xv), C.byref(gv), dw9.data_ptr(), 1 if i else 0, st.scratch.data_ptr(), st.scratch.numel() * 4, _stream(dev)))
        return (None, None, dw9.t().reshape(Cc, 1, 3, 3)) + tuple(dxs)


class WeightedSumSwish(Function):
    """bifpn.py:170-231: swish(sum_i w_i * in_i) with the (already normalised) fusion weights ``wn`` on the device."""

    @staticmethod
    def forward(ctx, st, wn, *ins):
        dev = ins[0].device
        s = torch.empty_like(ins[0], memory_format=torch.contiguous_format)
        a = torch.empty_like(s)
        wn32 = wn.detach().float().contiguous()
        d = nv.WsumDesc()
        d.n_in = len(ins)
        for k, t in enumerate(ins):
            d.in_[k] = _mat(t)
        d.w, d.s, d.a = wn32.data_ptr(), _mat(s), _mat(a)
        nv.check(nv.lib.hn_wsum_swish_fwd(C.byref(d), _stream(dev)))
        ctx.st = st
        ctx.save_for_backward(s, wn32, *ins)
        return a

    @staticmethod
    def backward(ctx, da):
        s, wn32, *ins = ctx.saved_tensors
        st = ctx.st
        dev = da.device
        dins = [torch.empty_like(s) for _ in ins]
        ds = _act_bwd(_c(da), s, nv.ACT_SWISH, scaled=dins, w=wn32)
        dots = torch.empty((len(ins), s.shape[-1]), dtype=torch.float32, device=dev)
        dsm = _mat(ds)
        for k, t in enumerate(ins):
            tm = _mat(t)
            nv.check(nv.lib.hn_col_reduce(C.byref(dsm), C.byref(tm), 1, 0, dots[k].data_ptr(), None, 1.0, st.scratch.data_ptr(), st.scratch.numel() * 4, _stream(dev)))
        return (None, dots.sum(dim=1)) + tuple(dins)


class Resample(Function):
    """Nearest x2 up-sampling / the two 3x3 stride-2 max-pools (common.py:117-151, lanedetect.py:41) and their adjoints."""

    @staticmethod
    def forward(ctx, x, mode):
        N, H, W, Cc = x.shape
        if mode == nv.RS_UP2:
            Ho, Wo = 2 * H, 2 * W
        elif mode == nv.RS_POOL_ZERO:
            Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
        else:
            Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, Ho, Wo, Cc), dtype=BF, device=x.device)
        d = nv.ResampleDesc(mode, _view(x), _view(y), _null_view(), _null_view())
        nv.check(nv.lib.hn_resample_fwd(C.byref(d), _stream(x.device)))
        ctx.mode = mode
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy if dy.stride(3) == 1 else dy.contiguous()
        dx = torch.empty(tuple(x.shape), dtype=BF, device=x.device)
        d = nv.ResampleDesc(ctx.mode, _view(x), _null_view(), _view(dy), _view(dx))
        nv.check(nv.lib.hn_resample_bwd(C.byref(d), _stream(x.device)))
        return dx, None


class HeadConv(Function):
    """Final 1x1 convolution of a head: bf16 rows in, fp32 tensor in the reference's output layout out (written by the GEMM
    epilogue), optional sigmoid.  ``layout`` = (out_shape, out_off, (stride_n, stride_pix), rows_per_img | None, groups | None):
    plain rows-per-image addressing (lane heads, lanedetect.py:84-96) or row groups = pyramid levels (detection.py:40-43)."""

    @staticmethod
    def forward(ctx, st, rec, act, layout, out, w, bias, x):
        dev = x.device
        shape, off, (sn, spix), rpi, groups = layout
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=dev)
        m = _mat(x)
        d = _conv_desc([_rows_view(x)], rec.fwd, flat=True, cout=rec.cout, out_ptr=out.data_ptr() + 4 * off, out_strides=(sn, 0, spix),
                       flat_hw=rpi if rpi else max(m.rows, 1), bias=bias.detach().data_ptr(), act=act, out_fp32=1, groups=groups)
        nv.check(nv.lib.hn_conv_fwd(C.byref(d), _stream(dev)))
        ctx.st, ctx.rec, ctx.act, ctx.layout, ctx.tok = st, rec, act, layout, st.token
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, out = ctx.saved_tensors
        st, rec = ctx.st, ctx.rec
        dev = x.device
        shape, off, (sn, spix), rpi, groups = ctx.layout
        dout = _c(dout)
        cp = (rec.cout + 7) // 8 * 8
        m = _mat(x)
        dz = torch.empty((m.rows, cp), dtype=BF, device=dev)
        hd = nv.HeadGradDesc()
        hd.dout, hd.out, hd.act, hd.cols_valid = dout.data_ptr() + 4 * off, out.data_ptr() + 4 * off, ctx.act, rec.cout
        hd.stride_n, hd.stride_pix, hd.stride_c, hd.rows_per_img = sn, spix, 1, rpi if rpi else 0
        if groups is not None:
            ends, hws, bases = groups
            hd.n_groups = len(ends)
            for i in range(len(ends)):
                hd.group_end[i], hd.group_hw[i], hd.group_out_base[i] = ends[i], hws[i], bases[i]
        hd.dz = _mat(dz)
        nv.check(nv.lib.hn_head_grad(C.byref(hd), _stream(dev)))
        dx = torch.empty_like(x, memory_format=torch.contiguous_format)
        _gemm_rows(dev, dz, rec.dgrad, rec.cin, dx)
        dw = ctx.tok.dw(rec)
        _wgrad(st, rec, dz, [x], _rows_view(dz), [_rows_view(x)], True, (1, 128), dw)
        db = _colsum(st, dz, rec.cout)
        return None, None, None, None, None, dw, db, dx


# --------------------------------------------------------------------------------------------------------------------
# the network (reference graph: anynet.py:136-145, bifpn.py:156-233, segmentation.py:84-105, detection.py:28-83,211-215,
# lanedetect.py:66-96, model.py:159-192)
# --------------------------------------------------------------------------------------------------------------------
def _bn(st, bn, z, act=nv.ACT_NONE, res=None):
    return BatchNormAct.apply(st, [bn], act, None, z, res, bn.weight, bn.bias)


def _xblock(st, p, blk, x):
    a = _bn(st, blk.conv_block_1[1], Conv1x1.apply(st, st.recs[p + ".c1"], None, blk.conv_block_1[0].weight, x), nv.ACT_RELU)
    g = _bn(st, blk.conv_block_2[1], GroupedConv3x3.apply(st, st.recs[p + ".c2"], blk.conv_block_2[0].weight, a), nv.ACT_RELU)
    if blk.se is not None:
        g = SqueezeExcite.apply(st, g, blk.se[1].weight, blk.se[1].bias, blk.se[3].weight, blk.se[3].bias)
    if blk.shortcut is not None:
        rec = st.recs[p + ".sc"]
        sc = Conv1x1S2.apply(st, rec, blk.shortcut[0].weight, x) if rec.kind == "pw_s2" else Conv1x1.apply(st, rec, None, blk.shortcut[0].weight, x)
        res = _bn(st, blk.shortcut[1], sc)
    else:
        res = x
    z = Conv1x1.apply(st, st.recs[p + ".c3"], None, blk.conv_block_3[0].weight, g)
    return _bn(st, blk.conv_block_3[1], z, nv.ACT_RELU, res)


def _backbone(st, m, x):
    net = m.backbone.net
    y = _bn(st, net.stem.bn, StemConv.apply(st, x, net.stem.conv.weight), nv.ACT_RELU)
    feats = []
    for s in range(m.backbone.stage_num):
        for bi, blk in enumerate(getattr(net, "stage_%d" % s).blocks.children()):
            y = _xblock(st, "backbone.s%d.b%d" % (s, bi), blk, y)
        feats.append(y)
    return feats


def _sepconv(st, name, sep, x):
    """SeparableConvBlock with norm (common.py:76-114): depthwise -> pointwise(+bias) -> BatchNorm."""
    t = Depthwise3x3.apply(st, False, sep.depthwise_conv.conv.weight, x)
    z = Conv1x1.apply(st, st.recs[name], sep.pointwise_conv.conv.bias, sep.pointwise_conv.conv.weight, t)
    return _bn(st, sep.bn, z)


def _reduce(st, name, red, x):
    return _bn(st, red[1], Conv1x1.apply(st, st.recs[name], red[0].conv.bias, red[0].conv.weight, x))


def _neck(st, m, feats):
    levels = None
    up = lambda t: Resample.apply(t, nv.RS_UP2)
    pool = lambda t: Resample.apply(t, nv.RS_POOL_ZERO)
    for ci, cell in enumerate(m.neck.bifpn.children()):
        p = "neck.c%d" % ci
        if cell.first_time:
            if len(feats) == 4:
                c3, c4, c5 = feats[-3:]
                p6_in = pool(_reduce(st, p + ".p5_to_p6", cell.p5_to_p6, c5))
            else:
                c3, c4, c5, c6 = feats[-4:]
                p6_in = _reduce(st, p + ".p6_down_channel", cell.p6_down_channel, c6)
            p7_in = pool(p6_in)
            p3_in = _reduce(st, p + ".p3_down_channel", cell.p3_down_channel, c3)
            p4_a = _reduce(st, p + ".p4_down_channel", cell.p4_down_channel, c4)
            p5_a = _reduce(st, p + ".p5_down_channel", cell.p5_down_channel, c5)
            p4_b = _reduce(st, p + ".p4_down_channel_2", cell.p4_down_channel_2, c4)
            p5_b = _reduce(st, p + ".p5_down_channel_2", cell.p5_down_channel_2, c5)
        else:
            p3_in, p4_a, p5_a, p6_in, p7_in = levels
            p4_b, p5_b = p4_a, p5_a

        def node(tag, wparam, ins):
            w = torch.relu(wparam)
            wn = w / (torch.sum(w, dim=0) + cell.epsilon)
            return _sepconv(st, p + "." + tag, getattr(cell, tag), WeightedSumSwish.apply(st, wn, *ins))

        p6_up = node("conv6_up", cell.p6_w1, [p6_in, up(p7_in)])
        p5_up = node("conv5_up", cell.p5_w1, [p5_a, up(p6_up)])
        p4_up = node("conv4_up", cell.p4_w1, [p4_a, up(p5_up)])
        p3_out = node("conv3_up", cell.p3_w1, [p3_in, up(p4_up)])
        p4_out = node("conv4_down", cell.p4_w2, [p4_b, p4_up, pool(p3_out)])
        p5_out = node("conv5_down", cell.p5_w2, [p5_b, p5_up, pool(p4_out)])
        p6_out = node("conv6_down", cell.p6_w2, [p6_in, p6_up, pool(p5_out)])
        p7_out = node("conv7_down", cell.p7_w2, [p7_in, pool(p6_out)])
        levels = [p3_out, p4_out, p5_out, p6_out, p7_out]
    return levels


def _seg_head(st, m, feats0, levels):
    dec = list(m.segheader.decoder.children())
    n = len(m.segheader.num_ch_enc)
    skips = [feats0] + list(levels[:n - 1])
    x = skips[-1]
    for i in range(n):
        c0, c1 = dec[2 * i].conv.conv, dec[2 * i + 1].conv.conv
        x = Conv3x3Padded.apply(st, st.recs["seg.d%d" % (2 * i)], nv.ACT_ELU, False, c0.weight, c0.bias, SegGather.apply(None, x))
        skip = skips[n - 2 - i] if i < n - 1 else None
        x = Conv3x3Padded.apply(st, st.recs["seg.d%d" % (2 * i + 1)], nv.ACT_ELU, False, c1.weight, c1.bias, SegGather.apply(x, skip))
    oc = dec[-1].conv
    logits = Conv3x3Padded.apply(st, st.recs["seg.d%d" % (2 * n)], nv.ACT_NONE, True, oc.weight, oc.bias, SegGather.apply(x, None))
    return logits.permute(0, 3, 1, 2)  # NCHW view of the NHWC logits (segmentation.py:105 returns [B, classes, H, W])


def _det_head(st, m, levels):
    dh = m.detectheader
    na, ncls = dh.num_anchors, dh.num_classes
    B = levels[0].shape[0]
    hws = [l.shape[1] * l.shape[2] for l in levels]
    total = sum(hws) * na
    ends, acc = [], 0
    for hw in hws:
        acc += B * hw
        ends.append(acc)
    shapes = [tuple(l.shape) for l in levels]
    outs = {}
    for tn, tower, k, act in (("reg", dh.regressor, 4, nv.ACT_NONE), ("cls", dh.classifier, ncls, nv.ACT_SIGMOID)):
        cur = list(levels)
        for i in range(tower.num_layers):
            sep = tower.conv_list[i]
            t = Depthwise3x3.apply(st, True, sep.depthwise_conv.conv.weight, *cur)
            z = Conv1x1.apply(st, st.recs["det.%s.%d" % (tn, i)], sep.pointwise_conv.conv.bias, sep.pointwise_conv.conv.weight, t)
            bns = [tower.bn_list[li][i] for li in range(len(levels))]
            y = BatchNormAct.apply(st, bns, nv.ACT_SWISH, ends, z, None, *([b.weight for b in bns] + [b.bias for b in bns]))
            cur = Depthwise3x3._level_views(y, shapes)
        sep = tower.header
        t = Depthwise3x3.apply(st, True, sep.depthwise_conv.conv.weight, *cur)
        cout = sep.pointwise_conv.conv.weight.shape[0]
        bases, a0 = [], 0
        for hw in hws:
            bases.append(a0 * k)
            a0 += hw * na
        layout = ((B, total, k), 0, (total * k, cout), None, (ends, hws, bases))
        outs[tn] = HeadConv.apply(st, st.recs["det.%s.hdr" % tn], act, layout, None, sep.pointwise_conv.conv.weight, sep.pointwise_conv.conv.bias, t)
    return outs["reg"], outs["cls"]


def _lane_head(st, m, levels):
    lh = m.laneheader
    if lh.stride != 32:
        raise NotImplementedError("train mode implements the lane head at anchor_stride 32 (the reference configs; lanedetect.py:76-80)")
    p3, p4, p5, p6 = levels[:4]
    mp = lambda t: Resample.apply(t, nv.RS_POOL_NEGINF)
    srcs = [mp(mp(p3)), mp(p4), p5, Resample.apply(p6, nv.RS_UP2)]
    B, fh, fw, _ = p5.shape
    ncls, nup, ndown = lh.num_classes, lh.lane_up_pts_num, lh.lane_down_pts_num
    pcls = torch.empty((B, fh * fw, ncls), dtype=torch.float32, device=p5.device)
    ploc_parts = {}
    for tag, br in (("cls", lh.conv_cls_conv), ("up", lh.conv_up_conv), ("down", lh.conv_down_conv)):
        hid = _bn(st, br[1], Conv1x1.apply(st, st.recs["lane.%s.hid" % tag], None, br[0].weight, *srcs), nv.ACT_RELU)
        cout = br[3].weight.shape[0]
        layout = ((B, fh * fw, cout), 0, (fh * fw * cout, cout), fh * fw, None)
        o = HeadConv.apply(st, st.recs["lane.%s.out" % tag], nv.ACT_NONE, layout, None, br[3].weight, br[3].bias, hid)
        if tag == "cls":
            pcls = o
        else:
            ploc_parts[tag] = o
    ploc = torch.cat([ploc_parts["down"], ploc_parts["up"]], dim=-1)  # lanedetect.py:93-95
    return pcls, ploc


def train_forward(model, x, mode="train"):
    """HydraNet.forward in train mode (model.py:159-192): returns the reference's output dict, autograd-connected."""
    dev = x.device
    st = get_state(model, dev)
    x = x.detach().float().contiguous()
    st.bn_seen = []
    st.token = StepToken(st)
    st.pack()
    feats = _backbone(st, model, x)
    levels = _neck(st, model, feats)
    out = {}
    if model.train_seg:
        out["seg"] = _seg_head(st, model, feats[0], levels)
    anchors = regression = classification = lane_cls = lane_reg = None
    if model.train_detect:
        regression, classification = _det_head(st, model, levels)
        anchors = model.detectheader.anchors(x, x.dtype)
        out["detection"] = {"anchors": anchors, "regression": regression, "classification": classification}
    if model.train_lane:
        lane_cls, lane_reg = _lane_head(st, model, levels)
        out["lane"] = dict(predict_cls=lane_cls, predict_loc=lane_reg)
    if st.bn_seen:  # nn.BatchNorm2d.forward also counts its calls
        torch._foreach_add_([b.num_batches_tracked for b in st.bn_seen], 1)
    if mode != "deploy":
        return out
    seg_cls = torch.argmax(out["seg"], dim=1) if model.train_seg else None
    return seg_cls, anchors, regression, classification, lane_cls, lane_reg


# --------------------------------------------------------------------------------------------------------------------
# the step as a callable, optionally replayed as ONE CUDA graph
# --------------------------------------------------------------------------------------------------------------------
def weighted_total(cfgs, loss_dict):
    """train.py:192-203: the yml's loss weights."""
    t = 0.0
    if "loss_seg" in loss_dict:
        t = t + loss_dict["loss_seg"] * cfgs["segment"].get("segment_weight", 1.0)
    if "loss_det_cls" in loss_dict:
        d = cfgs["detection"]
        t = t + (loss_dict["loss_det_cls"] * d.get("loss_cls_weight", 1.0) + loss_dict["loss_det_reg"] * d.get("loss_reg_weight", 1.0)) * d.get("detection_weight", 1.0)
    if "loss_lane_cls_pos" in loss_dict:
        l = cfgs["lane"]
        t = t + (loss_dict["loss_lane_cls_pos"] * l.get("loss_cls_pos_weight", 1.0) + loss_dict["loss_lane_cls_neg"] * l.get("loss_cls_neg_weight", 1.0)
                 + loss_dict["loss_lane_loc"] * l.get("loss_loc_weight", 1.0)) * l.get("lane_weight", 1.0)
    return t


class TrainStep:
    """One iteration of the reference's training loop (train.py:246-267) as a callable::

        step = TrainStep(hydranet, optimizer)          # optimizer: FusedAdam or any torch.optim.Optimizer
        loss = step(inputs, batch)                      # forward -> cal_loss -> weighted total -> zero_grad -> backward -> (all-reduce) -> step

    ``graph=True`` captures forward + loss + backward (and a FusedAdam step when there is no gradient exchange) into ONE CUDA
    graph on the first call and replays it afterwards: the step launches ~5 000 small kernels, which the Python / launch path
    cannot feed as fast as the GPU retires them.  Inputs are copied into static buffers; the loss losses are sync-free (losses.py),
    so nothing in the step waits for the host.  With several ranks (``torch.distributed`` initialised) the gradients are
    averaged after the graph by one coalesced NCCL all-reduce over all gradient tensors, in place.  In eager mode
    (``graph=False``) pass a ``GradAllReduce`` as ``reducer`` to overlap the bucketed exchange with backward instead.
    """

    def __init__(self, model, optimizer, graph=True, reducer=None, warmup=2, process_group=None, overlap_exchange=True):
        import torch.distributed as dist
        self.model, self.opt, self.use_graph, self.reducer, self.warmup = model, optimizer, graph, reducer, warmup
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1 and graph and reducer is None and overlap_exchange and os.environ.get("HN_GRAPH_OVERLAP", "1") != "0":
            # the bucketed exchange INSIDE the captured step: the hooks fire during the captured backward, so the NCCL kernels
            # become nodes of the graph on a forked stream and overlap the rest of backward (NCCL collectives are capturable)
            from .parallel import GradAllReduce
            self.reducer = GradAllReduce(model.parameters(), process_group=process_group, model=model)
        self.graph = None
        self.static = None
        self.opt_in_graph = False
        self.loss_dict = None

    def _fwd_bwd(self, x, gt):
        out = self.model(x)
        ld = self.model.cal_loss(out, gt)
        loss = weighted_total(self.model.cfgs, ld)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        st = getattr(self.model, "_train_state", None)
        if st is not None:
            join_side_stream(st, x.device)
        if self.reducer is not None:
            self.reducer.finish()
        return loss, ld

    def _exchange(self):
        if self.world <= 1 or self.reducer is not None:
            return
        import torch.distributed as dist
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        if grads[0].is_cuda:
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
        else:
            for g in grads:
                dist.all_reduce(g, group=self.group)
                g.div_(self.world)

    def _eager(self, x, gt):
        loss, ld = self._fwd_bwd(x, gt)
        self._exchange()
        self.opt.step()
        self.loss_dict = ld
        return loss.detach()

    def _capture(self, x, gt):
        from .optim import FusedAdam
        dev = x.device
        self.static = (torch.empty_like(x), {k: torch.empty_like(v) for k, v in gt.items()})
        self.static[0].copy_(x)
        for k, v in gt.items():
            self.static[1][k].copy_(v)
        # the warm-up iterations and the capture must not change the training trajectory: snapshot parameters, buffers and
        # optimizer state, restore them (in place: the graph holds their addresses) afterwards
        named = list(self.model.parameters()) + list(self.model.buffers())
        saved = [t.detach().clone() for t in named]
        pre = {}
        for st in self.opt.state.values():
            for k, v in st.items():
                pre[(id(st), k)] = v.detach().clone() if torch.is_tensor(v) else v
        pre_dyn = {gi: t["dyn"].clone() for gi, t in getattr(self.opt, "_tables", {}).items()}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, self.warmup)):
                self._fwd_bwd(*self.static)
                self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.opt_in_graph = isinstance(self.opt, FusedAdam) and self.world <= 1
        self.opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss, ld = self._fwd_bwd(*self.static)
            if self.opt_in_graph:
                self.opt.step()
        self.graph, self.static_loss, self.loss_dict = g, loss.detach(), {k: v.detach() for k, v in ld.items()}
        with torch.no_grad():
            for t, s in zip(named, saved):
                t.copy_(s)
            for st in self.opt.state.values():  # optimizer state back to its pre-warm-up values (zeros / step 0 if it was fresh)
                for k, v in list(st.items()):
                    old = pre.get((id(st), k))
                    if torch.is_tensor(v):
                        v.copy_(old) if old is not None else v.zero_()
                    elif k == "step":
                        st[k] = old if old is not None else 0
            for gi, t in getattr(self.opt, "_tables", {}).items():
                t["dyn"][1:2].copy_(pre_dyn[gi][1:2]) if gi in pre_dyn else t["dyn"][1:2].zero_()

    def close(self):
        """Release the captured graph (and the hooks of an internal reducer).  Call before ``destroy_process_group()`` when the
        graph holds captured NCCL collectives: the communicator cannot shut down while a live graph still references it."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
        if self.reducer is not None:
            self.reducer.remove()

    def __call__(self, x, gt):
        if not self.use_graph:
            return self._eager(x, gt)
        if self.graph is None:
            self._capture(x, gt)
        self.static[0].copy_(x, non_blocking=True)
        for k, v in gt.items():
            self.static[1][k].copy_(v, non_blocking=True)
        if hasattr(self.opt, "sync_hyper"):
            self.opt.sync_hyper()
        self.graph.replay()
        if not self.opt_in_graph:
            self._exchange()
            self.opt.step()
        else:
            for st in self.opt.state.values():
                if "step" in st and not torch.is_tensor(st["step"]):
                    st["step"] += 1
        return self.static_loss
