"""Parameter tree of HydraNet.

These classes only *hold* parameters/buffers under exactly the names the reference uses, so a
reference ``state_dict`` loads unchanged (1177 entries for the big cfg).
They contain no arithmetic: the eval-mode forward is executed by the native engine
(``engine.py`` -> ``csrc/``); nothing here calls ``nn.Conv2d.forward``.

Name parity with the reference (file:line in /root/reference/model):
  backbone.net.stem / stage_i.blocks.block_j.{conv_block_1,conv_block_2,se,conv_block_3,shortcut}
      net/anynet.py:8-90, width derivation net/regnet.py:22-44
  neck.bifpn.k.{convN_up,convN_down,pN_down_channel,...,pN_w1,pN_w2}   net/bifpn.py:12-123
  segheader.decoder.i.conv.conv / decoder.8.conv                        head_seg/segmentation.py:51-82
  detectheader.{regressor,classifier}.{conv_list,bn_list,header}        head_detect/detection.py:11-83
  laneheader.conv_{cls,up,down}_conv                                    head_lane/lanedetect.py:45-64
"""
import math

import numpy as np
import torch
from torch import nn


def regnet_stage_plan(initial_width, slope, quantized_param, network_depth, bottleneck_ratio, group_width):
    """Per-stage (num_blocks, width, group_width) -- same arithmetic as net/regnet.py:22-39."""
    u = initial_width + slope * np.arange(network_depth)
    k = np.round(np.log(u / initial_width) / np.log(quantized_param))
    w = initial_width * np.power(quantized_param, k)
    w = 8 * np.round(w / 8)
    widths, counts = np.unique(w.astype(np.int32), return_counts=True)
    gws = np.array([min(group_width, bw // bottleneck_ratio) for bw in widths])
    widths = np.round(widths // bottleneck_ratio / group_width) * group_width
    gws = gws.astype(np.int32) * bottleneck_ratio
    return [(int(n), int(wd), int(g)) for n, wd, g in zip(counts, widths.astype(np.int32), gws)]


def _seq(*mods):
    return nn.Sequential(*mods)


class _Stem(nn.Module):
    def __init__(self, cout):
        super().__init__()
        self.conv = nn.Conv2d(3, cout, 3, stride=2, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(cout)


class _XBlock(nn.Module):
    def __init__(self, cin, cout, bott, gw, stride, se_ratio):
        super().__init__()
        mid = cout // bott
        self.cin, self.cout, self.mid, self.stride, self.groups = cin, cout, mid, stride, mid // gw
        self.conv_block_1 = _seq(nn.Conv2d(cin, mid, 1, bias=False), nn.BatchNorm2d(mid), nn.ReLU())
        self.conv_block_2 = _seq(nn.Conv2d(mid, mid, 3, stride=stride, groups=mid // gw, padding=1, bias=False),
                                 nn.BatchNorm2d(mid), nn.ReLU())
        if se_ratio is not None:
            sc = cin // se_ratio  # SE width follows the block *input* channels (anynet.py:41)
            self.se = _seq(nn.AdaptiveAvgPool2d(1), nn.Conv2d(mid, sc, 1), nn.ReLU(), nn.Conv2d(sc, mid, 1), nn.Sigmoid())
        else:
            self.se = None
        self.conv_block_3 = _seq(nn.Conv2d(mid, cout, 1, bias=False), nn.BatchNorm2d(cout))
        if stride != 1 or cin != cout:
            self.shortcut = _seq(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))
        else:
            self.shortcut = None


class _Stage(nn.Module):
    def __init__(self, n, cin, cout, bott, gw, stride, se_ratio):
        super().__init__()
        self.blocks = nn.Sequential()
        for i in range(n):
            self.blocks.add_module("block_%d" % i,
                                   _XBlock(cin if i == 0 else cout, cout, bott, gw, stride if i == 0 else 1, se_ratio))


class RegNetY(nn.Module):
    def __init__(self, initial_width, slope, quantized_param, network_depth, bottleneck_ratio, group_width, stride, se_ratio):
        super().__init__()
        plan = regnet_stage_plan(initial_width, slope, quantized_param, network_depth, bottleneck_ratio, group_width)
        self.net = nn.Sequential()
        self.net.add_module("stem", _Stem(32))
        prev = 32
        for i, (n, w, g) in enumerate(plan):
            assert w % (bottleneck_ratio * g) == 0
            self.net.add_module("stage_%d" % i, _Stage(n, prev, w, bottleneck_ratio, g, stride, se_ratio))
            prev = w
        self.stage_num = len(plan)
        self.stage_widths = [w for _, w, _ in plan]
        for m in self.modules():  # init as anynet.py:124-134
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0.0, math.sqrt(2.0 / fan_out))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1.0)
                m.bias.data.zero_()


class _SameConv(nn.Module):
    """Holder for net/common.py:35-73 (params live under ``.conv``)."""

    def __init__(self, cin, cout, k, bias=True, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, bias=bias, groups=groups)


class _SepConv(nn.Module):
    """Holder for SeparableConvBlock net/common.py:76-114."""

    def __init__(self, cin, cout=None, norm=True):
        super().__init__()
        cout = cin if cout is None else cout
        self.depthwise_conv = _SameConv(cin, cin, 3, bias=False, groups=cin)
        self.pointwise_conv = _SameConv(cin, cout, 1)
        self.norm = norm
        if norm:
            self.bn = nn.BatchNorm2d(cout, momentum=0.01, eps=1e-3)


def _reducer(cin, cout, pool=False):
    mods = [_SameConv(cin, cout, 1), nn.BatchNorm2d(cout, momentum=0.01, eps=1e-3)]
    if pool:
        mods.append(nn.Identity())  # MaxPool2dStaticSamePadding has no parameters
    return _seq(*mods)


class _BiFPN(nn.Module):
    NODES = ("conv6_up", "conv5_up", "conv4_up", "conv3_up", "conv4_down", "conv5_down", "conv6_down", "conv7_down")

    def __init__(self, ch, conv_channels, first_time, epsilon=1e-4):
        super().__init__()
        self.epsilon, self.first_time = epsilon, first_time
        for n in self.NODES:
            setattr(self, n, _SepConv(ch))
        if first_time:
            self.p5_down_channel = _reducer(conv_channels[2], ch)
            self.p4_down_channel = _reducer(conv_channels[1], ch)
            self.p3_down_channel = _reducer(conv_channels[0], ch)
            self.p5_to_p6 = _reducer(conv_channels[2], ch, pool=True)
            if len(conv_channels) == 4:
                self.p6_down_channel = _reducer(conv_channels[3], ch)
            self.p4_down_channel_2 = _reducer(conv_channels[1], ch)
            self.p5_down_channel_2 = _reducer(conv_channels[2], ch)
        for n in ("p6_w1", "p5_w1", "p4_w1", "p3_w1"):
            setattr(self, n, nn.Parameter(torch.ones(2)))
        self.p4_w2 = nn.Parameter(torch.ones(3))
        self.p5_w2 = nn.Parameter(torch.ones(3))
        self.p6_w2 = nn.Parameter(torch.ones(3))
        self.p7_w2 = nn.Parameter(torch.ones(2))


class StackBiFPN(nn.Module):
    def __init__(self, fpn_num_filters, fpn_cell_repeats, conv_channel_coef):
        super().__init__()
        self.bifpn = nn.Sequential(*[_BiFPN(fpn_num_filters, conv_channel_coef, i == 0) for i in range(fpn_cell_repeats)])


class _Conv3x3(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(int(cin), int(cout), 3)


class _ConvBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _Conv3x3(cin, cout)


class _Tower(nn.Module):
    def __init__(self, ch, cout, num_layers, pyramid_levels):
        super().__init__()
        self.num_layers = num_layers
        self.conv_list = nn.ModuleList([_SepConv(ch, ch, norm=False) for _ in range(num_layers)])
        self.bn_list = nn.ModuleList([nn.ModuleList([nn.BatchNorm2d(ch, momentum=0.01, eps=1e-3) for _ in range(num_layers)])
                                      for _ in range(pyramid_levels)])
        self.header = _SepConv(ch, cout, norm=False)
