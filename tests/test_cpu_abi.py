"""CPU suite: the C-ABI library loads without a GPU, exports every symbol include/hydranet_b200.h
declares, and the ctypes mirrors of the descriptor structs have the C compiler's layout."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

import hydranet_b200  # noqa: F401
from hydranet_b200 import _native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hydranet_b200.h")


def declared_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^\s*(?:const\s+)?[A-Za-z_0-9]+\s*\*?\s*(hn_[a-z0-9_]+)\s*\(", src, re.M)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(nv.lib, n), "not exported: " + n
        assert n in nv.SYMBOLS, "no ctypes prototype: " + n
    assert sorted(nv.SYMBOLS) == names
    assert nv.lib.hn_version() >= 100
    assert isinstance(nv.lib.hn_last_error(), bytes)


def test_struct_layouts_match_the_c_compiler():
    structs = {"hn_view": nv.View, "hn_tap": nv.Tap, "hn_conv_desc": nv.ConvDesc, "hn_stem_desc": nv.StemDesc,
               "hn_node_desc": nv.NodeDesc, "hn_dw_multi_desc": nv.DwMultiDesc, "hn_pool_desc": nv.PoolDesc, "hn_lanefuse_desc": nv.LaneFuseDesc,
               "hn_se_pool_desc": nv.SePoolDesc, "hn_se_scale_desc": nv.SeScaleDesc, "hn_gconv_se_desc": nv.GconvSeDesc, "hn_segloss_desc": nv.SegLossDesc, "hn_det_desc": nv.DetDesc, "hn_lane_desc": nv.LaneDesc,
               "hn_mat": nv.Mat, "hn_bn_desc": nv.BnDesc, "hn_actbwd_desc": nv.ActBwdDesc, "hn_wsum_desc": nv.WsumDesc,
               "hn_resample_desc": nv.ResampleDesc, "hn_seggather_desc": nv.SegGatherDesc, "hn_headgrad_desc": nv.HeadGradDesc,
               "hn_sefc_desc": nv.SeFcDesc, "hn_pack_entry": nv.PackEntry, "hn_wgrad_desc": nv.WgradDesc, "hn_adam_tensor": nv.AdamTensor}
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "hydranet_b200.h"\nint main(){\n'
    for c in structs:
        prog += 'printf("%s %%zu\\n", sizeof(%s));\n' % (c, c)
    prog += 'printf("conv.out %zu\\n", offsetof(hn_conv_desc, out));\nprintf("conv.n_cls %zu\\n", offsetof(hn_conv_desc, n_cls));\n'
    prog += 'printf("lane.out_x %zu\\n", offsetof(hn_lane_desc, out_x));\nprintf("det.pre_boxes %zu\\n", offsetof(hn_det_desc, pre_boxes));\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = dict(l.split() for l in subprocess.check_output([os.path.join(d, "t")], text=True).splitlines())
    for c, py in structs.items():
        assert int(out[c]) == ctypes.sizeof(py), c
    assert int(out["conv.out"]) == nv.ConvDesc.out.offset and int(out["conv.n_cls"]) == nv.ConvDesc.n_cls.offset
    assert int(out["lane.out_x"]) == nv.LaneDesc.out_x.offset and int(out["det.pre_boxes"]) == nv.DetDesc.pre_boxes.offset


def test_argument_errors_are_reported_not_crashed():
    d = nv.ConvDesc()
    assert nv.lib.hn_conv_fwd(ctypes.byref(d), None) == 1  # HN_ERR_ARG, before any CUDA call
    assert b"n_src" in nv.lib.hn_last_error()
    with pytest.raises(nv.NativeError):
        nv.check(nv.lib.hn_seg_argmax(None, 1, 5, 16, None, None, None))
    assert nv.lib.hn_plan_run(None, None) == 1


def test_no_cpu_fallback():
    import torch
    from hydranet_b200.config import big_cfg
    m = hydranet_b200.HydraNet(big_cfg(128, 128)).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 128, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hydranet_b200.SegmentHeader.argmax(torch.zeros(1, 5, 8, 8))
