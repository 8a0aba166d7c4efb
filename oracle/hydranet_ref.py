"""ORACLE -- test infrastructure only (never imported by the product path).

Functional fp32 restatement of the reference's ``HydraNet.forward`` on plain ``torch.nn.functional``
ops, driven by a reference-format ``state_dict``.  It exists so that the parity checker can travel to
the GPU box, where /root/reference does not exist.  Pinned bit-exactly against the live reference
in this container by ``tests/test_oracle_pinning.py`` / ``oracle/make_golden.py``.

Reference call sites restated (all under /root/reference/model):
  backbone   net/anynet.py:8-20 (Stem), 64-76 (XBlock.forward), 136-145 (AnyNetX.forward)
  neck       net/bifpn.py:156-233 (_forward_fast_attention), net/common.py:76-151
  seg head   head_seg/segmentation.py:84-105
  detect     head_detect/detection.py:28-44, 65-83, 108-170, 211-215
  lane       head_lane/lanedetect.py:66-96
  facade     model.py:159-198
"""
import contextlib
import itertools

import numpy as np
import torch
import torch.nn.functional as F


@contextlib.contextmanager
def ieee_fp32():
    """True fp32 on CUDA: torch's defaults let cuDNN convolutions (and, if enabled, matmuls) run on TF32 tensor cores
    (10-bit mantissa) -- useless as a parity oracle for a bf16 engine.  Every oracle forward runs inside this guard."""
    conv, mm = torch.backends.cudnn.conv.fp32_precision, torch.backends.cuda.matmul.fp32_precision
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    torch.backends.cuda.matmul.fp32_precision = "ieee"
    try:
        yield
    finally:
        torch.backends.cudnn.conv.fp32_precision, torch.backends.cuda.matmul.fp32_precision = conv, mm


_TRAIN = False  # set by forward(train=True): batch statistics + running-stat update, as nn.BatchNorm2d in .train() mode


def _bn(sd, p, x, eps):
    # momentum follows the reference's constructors: default 0.1 with eps 1e-5 (backbone, lane head), 0.01 with eps 1e-3
    # (neck, detection head: common.py:97, detection.py:22)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], _TRAIN,
                        (0.1 if eps == 1e-5 else 0.01) if _TRAIN else 0.0, eps)


def _swish(x):
    return x * torch.sigmoid(x)


def _xblock(sd, p, x, stride, groups):
    y = F.relu(_bn(sd, p + ".conv_block_1.1", F.conv2d(x, sd[p + ".conv_block_1.0.weight"]), 1e-5))
    y = F.conv2d(y, sd[p + ".conv_block_2.0.weight"], None, stride, 1, 1, groups)
    y = F.relu(_bn(sd, p + ".conv_block_2.1", y, 1e-5))
    if (p + ".se.1.weight") in sd:
        s = F.adaptive_avg_pool2d(y, 1)
        s = F.relu(F.conv2d(s, sd[p + ".se.1.weight"], sd[p + ".se.1.bias"]))
        s = torch.sigmoid(F.conv2d(s, sd[p + ".se.3.weight"], sd[p + ".se.3.bias"]))
        y = y * s
    y = _bn(sd, p + ".conv_block_3.1", F.conv2d(y, sd[p + ".conv_block_3.0.weight"]), 1e-5)
    if (p + ".shortcut.0.weight") in sd:
        x = _bn(sd, p + ".shortcut.1", F.conv2d(x, sd[p + ".shortcut.0.weight"], None, stride), 1e-5)
    return F.relu(y + x)


def backbone(sd, x, group_width=8, stride=2):
    x = F.relu(_bn(sd, "backbone.net.stem.bn", F.conv2d(x, sd["backbone.net.stem.conv.weight"], None, 2, 1), 1e-5))
    feats, s = [], 0
    while ("backbone.net.stage_%d.blocks.block_0.conv_block_1.0.weight" % s) in sd:
        b = 0
        while True:
            p = "backbone.net.stage_%d.blocks.block_%d" % (s, b)
            if (p + ".conv_block_1.0.weight") not in sd:
                break
            w2 = sd[p + ".conv_block_2.0.weight"]
            x = _xblock(sd, p, x, stride if b == 0 else 1, w2.shape[0] // w2.shape[1])
            b += 1
        feats.append(x)
        s += 1
    return feats


def _sepconv(sd, p, x, norm=True):
    c = x.shape[1]
    x = F.conv2d(F.pad(x, [1, 1, 1, 1]), sd[p + ".depthwise_conv.conv.weight"], None, 1, 0, 1, c)
    x = F.conv2d(x, sd[p + ".pointwise_conv.conv.weight"], sd[p + ".pointwise_conv.conv.bias"])
    if norm:
        x = _bn(sd, p + ".bn", x, 1e-3)
    return x


def _pool_same(x):  # MaxPool2dStaticSamePadding(3, 2): zeros padded right/bottom take part in the max
    return F.max_pool2d(F.pad(x, [0, 1, 0, 1]), 3, 2)


def _reduce(sd, p, x):
    return _bn(sd, p + ".1", F.conv2d(x, sd[p + ".0.conv.weight"], sd[p + ".0.conv.bias"]), 1e-3)


def _up(x, f=2):
    return F.interpolate(x, scale_factor=f, mode="nearest")


def bifpn_cell(sd, p, inputs, first_time, eps=1e-4):
    if first_time:
        if len(inputs) == 4:
            p3, p4, p5 = inputs[-3:]
            p6_in = _pool_same(_reduce(sd, p + ".p5_to_p6", p5))
        else:
            p3, p4, p5, p6r = inputs[-4:]
            p6_in = _reduce(sd, p + ".p6_down_channel", p6r)
        p7_in = _pool_same(p6_in)
        p3_in = _reduce(sd, p + ".p3_down_channel", p3)
        p4_in = _reduce(sd, p + ".p4_down_channel", p4)
        p5_in = _reduce(sd, p + ".p5_down_channel", p5)
    else:
        p4 = p5 = None
        p3_in, p4_in, p5_in, p6_in, p7_in = inputs

    def wt(name):
        w = F.relu(sd[p + "." + name])
        return w / (torch.sum(w, dim=0) + eps)

    w = wt("p6_w1")
    p6_up = _sepconv(sd, p + ".conv6_up", _swish(w[0] * p6_in + w[1] * _up(p7_in)))
    w = wt("p5_w1")
    p5_up = _sepconv(sd, p + ".conv5_up", _swish(w[0] * p5_in + w[1] * _up(p6_up)))
    w = wt("p4_w1")
    p4_up = _sepconv(sd, p + ".conv4_up", _swish(w[0] * p4_in + w[1] * _up(p5_up)))
    w = wt("p3_w1")
    p3_out = _sepconv(sd, p + ".conv3_up", _swish(w[0] * p3_in + w[1] * _up(p4_up)))
    if first_time:
        p4_in = _reduce(sd, p + ".p4_down_channel_2", p4)
        p5_in = _reduce(sd, p + ".p5_down_channel_2", p5)
    w = wt("p4_w2")
    p4_out = _sepconv(sd, p + ".conv4_down", _swish(w[0] * p4_in + w[1] * p4_up + w[2] * _pool_same(p3_out)))
    w = wt("p5_w2")
    p5_out = _sepconv(sd, p + ".conv5_down", _swish(w[0] * p5_in + w[1] * p5_up + w[2] * _pool_same(p4_out)))
    w = wt("p6_w2")
    p6_out = _sepconv(sd, p + ".conv6_down", _swish(w[0] * p6_in + w[1] * p6_up + w[2] * _pool_same(p5_out)))
    w = wt("p7_w2")
    p7_out = _sepconv(sd, p + ".conv7_down", _swish(w[0] * p7_in + w[1] * _pool_same(p6_out)))
    return p3_out, p4_out, p5_out, p6_out, p7_out


def neck(sd, feats):
    x, i = feats, 0
    while ("neck.bifpn.%d.p6_w1" % i) in sd:
        x = bifpn_cell(sd, "neck.bifpn.%d" % i, x, i == 0)
        i += 1
    return x


def _conv3x3_reflect(sd, p, x):
    return F.conv2d(F.pad(x, [1, 1, 1, 1], mode="reflect"), sd[p + ".weight"], sd[p + ".bias"])


def seg_head(sd, feats):
    n = len(feats)
    x = feats[-1]
    for i in range(n):
        x = F.elu(_conv3x3_reflect(sd, "segheader.decoder.%d.conv.conv" % (2 * i), x))
        xs = [_up(x)]
        if i < n - 1:
            xs.append(feats[n - 2 - i])
        x = torch.cat(xs, 1)
        x = F.elu(_conv3x3_reflect(sd, "segheader.decoder.%d.conv.conv" % (2 * i + 1), x))
    return _conv3x3_reflect(sd, "segheader.decoder.%d.conv" % (2 * n), _up(x))


def _tower(sd, p, inputs, num_layers, k):
    outs = []
    for li, feat in enumerate(inputs):
        for i in range(num_layers):
            feat = _sepconv(sd, "%s.conv_list.%d" % (p, i), feat, norm=False)
            feat = _swish(_bn(sd, "%s.bn_list.%d.%d" % (p, li, i), feat, 1e-3))
        feat = _sepconv(sd, p + ".header", feat, norm=False)
        feat = feat.permute(0, 2, 3, 1).contiguous()
        outs.append(feat.view(feat.shape[0], -1, k))
    return torch.cat(outs, 1)


def anchors(image_hw, anchor_scale, pyramid_levels, scales, ratios):
    H, W = image_hw
    boxes_all = []
    for stride in [2 ** l for l in pyramid_levels]:
        boxes_level = []
        for scale, ratio in itertools.product(scales, ratios):
            if W % stride != 0 or H % stride != 0:
                raise ValueError('input size must be divided by the stride.')
            base = anchor_scale * stride * scale
            ax2, ay2 = base * ratio[0] / 2.0, base * ratio[1] / 2.0
            xv, yv = np.meshgrid(np.arange(stride / 2, W, stride), np.arange(stride / 2, H, stride))
            xv, yv = xv.reshape(-1), yv.reshape(-1)
            boxes = np.swapaxes(np.vstack((yv - ay2, xv - ax2, yv + ay2, xv + ax2)), 0, 1)
            boxes_level.append(np.expand_dims(boxes, axis=1))
        boxes_all.append(np.concatenate(boxes_level, axis=1).reshape([-1, 4]))
    return torch.from_numpy(np.vstack(boxes_all).astype(np.float32)).unsqueeze(0)


def lane_head(sd, fused, stride, num_classes, n_loc):
    mp = lambda t: F.max_pool2d(t, 3, 2, 1)
    if stride == 16:
        x = torch.cat([mp(fused[0]), _up(fused[2]), fused[1], _up(fused[3], 4)], 1)
    elif stride == 32:
        x = torch.cat([mp(mp(fused[0])), mp(fused[1]), fused[2], _up(fused[3])], 1)
    else:
        raise ValueError("unsupported lane stride")

    def branch(p):
        y = F.relu(_bn(sd, p + ".1", F.conv2d(x, sd[p + ".0.weight"]), 1e-5))
        return F.conv2d(y, sd[p + ".3.weight"], sd[p + ".3.bias"]).permute(0, 2, 3, 1)

    cls = branch("laneheader.conv_cls_conv").contiguous()
    cls = cls.view(cls.shape[0], -1, num_classes)
    loc = torch.cat([branch("laneheader.conv_down_conv"), branch("laneheader.conv_up_conv")], -1).contiguous()
    return cls, loc.view(loc.shape[0], -1, n_loc)


def forward(sd, cfg, x, want_feats=False, train=False):
    """state_dict + cfg + fp32 NCHW input -> the reference's output dict (model.py:159-192).  IEEE fp32 on any device.
    train=True: BatchNorm uses batch statistics and updates sd's running statistics in place (train.py:242); tensors of
    ``sd`` that require grad stay leaves, so ``backward`` on a loss of the outputs yields the reference's gradients."""
    global _TRAIN
    _TRAIN = bool(train)
    try:
        with ieee_fp32():
            return _forward(sd, cfg, x, want_feats)
    finally:
        _TRAIN = False


def _forward(sd, cfg, x, want_feats=False):
    sd = {k: v.to(x.device) for k, v in sd.items()}
    feats = backbone(sd, x, cfg["backbone"]["group_width"], cfg["backbone"]["stride"])
    fused = neck(sd, feats)
    out = {}
    if cfg["train"]["train_seg"]:
        out["seg"] = seg_head(sd, [feats[0], fused[0], fused[1], fused[2]])
    if cfg["train"]["train_detect"]:
        dc = cfg["detection"]
        r1, r2 = dc["aspect_ratios_factor"]
        ratios = [(1.0, 1.0), (r1, r2), (r2, r1)]
        scales = [2 ** s for s in dc["scales_factor"]]
        na = len(ratios) * len(scales)
        levels = list(range(3, 3 + dc["pyramid_levels"]))
        out["detection"] = {
            "anchors": anchors(x.shape[2:], dc["anchor_scale"], levels, scales, ratios).to(x.device),
            "regression": _tower(sd, "detectheader.regressor", fused, dc["box_class_repeats"], 4),
            "classification": _tower(sd, "detectheader.classifier", fused, dc["box_class_repeats"], dc["num_classes"]).sigmoid(),
        }
    if cfg["train"]["train_lane"]:
        lc = cfg["lane"]
        ppl = int(cfg["dataloader"]["network_input_height"] / lc["interval"])
        cls, loc = lane_head(sd, fused, lc["anchor_stride"], lc["num_classes"], 2 * (ppl + 1))
        out["lane"] = dict(predict_cls=cls, predict_loc=loc)
    if want_feats:
        out["_feats"], out["_fused"] = feats, fused
    return out
