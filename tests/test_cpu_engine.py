"""CPU suite: host-side logic of the engine.  The op list the native plan would execute is run by the
pure-torch interpreter in tests/emulator.py (kernel semantics, fp32) and compared with the oracle:
this covers BN folding, weight packing, tap tables, views, halo logic and head wiring without a GPU."""
import os

import numpy as np
import pytest
import torch

import emulator
import hydranet_b200 as hb
from hydranet_b200 import engine
from hydranet_b200.config import big_cfg, small_cfg
from hydranet_b200.modules import regnet_stage_plan
from oracle import hydranet_ref, ref_live, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,cfg,hw", [("big", big_cfg(128, 128), (128, 128)), ("small", small_cfg(256, 128), (128, 256)),
                                          ("big_nonsquare", big_cfg(384, 256), (256, 384))])
def test_schedule_equals_oracle(name, cfg, hw):
    m = hb.HydraNet(cfg).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=2, seg_logit_gain=20.0)
    m.load_state_dict(sd)
    x = synth.synth_input(2, hw[0], hw[1], seed=4)
    with torch.no_grad():
        ref = hydranet_ref.forward(sd, cfg, x)
        b = engine.Builder(m, 2, hw[0], hw[1], torch.device("cpu"), act_dtype=torch.float32).build(x)
        emulator.run_ops(b.ops)
    for k, r, t in (("seg", ref["seg"], b.out["seg"]), ("reg", ref["detection"]["regression"], b.out["regression"]),
                    ("cls", ref["detection"]["classification"], b.out["classification"]),
                    ("lane_cls", ref["lane"]["predict_cls"], b.out["predict_cls"]), ("lane_loc", ref["lane"]["predict_loc"], b.out["predict_loc"])):
        assert float((r - t).abs().max()) <= 2e-5 * float(r.abs().max()), k
    assert float((torch.argmax(ref["seg"], 1) == b.out["seg_cls_u8"].long()).float().mean()) > 0.9999
    for op in b.ops:
        if op.kind == "conv":
            assert len(op.taps) <= hb._native.HN_MAX_TAPS and len(op.src) <= hb._native.HN_MAX_SRC
            assert op.weight.shape[0] % op.bn == 0 and op.bn % 16 == 0 and 16 <= op.bn <= 256
            assert 1024 + op.stages * (16384 + op.bn * 128) <= 227 * 1024


def test_heads_optional():
    cfg = big_cfg(128, 128)
    cfg["train"].update(train_seg=False, train_lane=False)
    m = hb.HydraNet(cfg).eval()
    assert m.segheader is None and m.laneheader is None and m.detectheader is not None
    x = synth.synth_input(1, 128, 128)
    b = engine.Builder(m, 1, 128, 128, torch.device("cpu"), act_dtype=torch.float32).build(x)
    emulator.run_ops(b.ops)
    assert set(b.out) == {"regression", "classification"}


def test_algorithmic_macs_match_survey():
    """Sum of per-op algorithmic MACs == the reference's conv MAC count."""
    m = hb.HydraNet(big_cfg()).eval()
    b = engine.Builder(m, 1, 640, 640, torch.device("cpu"), act_dtype=torch.float32)
    b.build(torch.zeros(1, 3, 640, 640))
    macs = sum(op.macs for op in b.ops if op.kind in ("conv", "stem", "node", "dw_multi", "se_pool", "se_fused", "gconv_se"))
    assert abs(macs - 31698401632) / 31698401632 < 2e-3, macs
    # launch structure of the big cfg: stride-1 XBlocks of stages 3-4 run their grouped 3x3 and squeeze-excite as
    # one launch, the other blocks of stages 2-4 their squeeze-excite; stages 0-1 keep pool + scale
    import collections
    kinds = collections.Counter(op.kind for op in b.ops)
    assert (kinds["conv"], kinds["gconv_se"], kinds["se_fused"], kinds["se_pool"], kinds["se_scale"]) == (134, 22, 6, 2, 2), kinds
    assert (kinds["node"], kinds["dw_multi"], kinds["stem"], kinds["pool"], kinds["lanefuse"]) == (24, 8, 1, 1, 1), kinds


def test_state_dict_keys_equal_reference():
    for name, cfg in (("big", big_cfg()), ("small", small_cfg())):
        want = [l.split(" ", 1) for l in open(os.path.join(GOLD, "state_dict_keys_%s.txt" % name)).read().splitlines()]
        sd = hb.HydraNet(cfg).state_dict()
        assert list(sd.keys()) == [k for k, _ in want]
        for k, rest in want:
            assert rest.startswith(str(tuple(sd[k].shape))), k


@pytest.mark.skipif(not ref_live.available(), reason="live reference only exists in the build container")
def test_reference_state_dict_loads_and_lane_host_tail():
    ref_model, RefCodec = ref_live.import_reference()
    import yaml
    cfg = yaml.safe_load(open("/root/reference/model/cfgs/hydranet_joint_big_backbone.yml"))
    torch.manual_seed(0)
    ref = ref_model.HydraNet(cfg).eval()
    mine = hb.HydraNet(cfg).eval()
    missing, unexpected = mine.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    # host tail after the lane kernel: order_lane_x_axis + convert_lane_to_dict (lane_codec_utils.py:185-282)
    from head_lane.lane_codec_utils import Lane as RL, Point as RP, order_lane_x_axis as r_order, convert_lane_to_dict as r_conv
    rng = np.random.default_rng(0)

    def mk(L, P):
        lanes = []
        for i in range(6):
            n = int(rng.integers(2, 9))
            xs = np.cumsum(rng.normal(0, 12, n)).astype(np.float32) + np.float32(rng.uniform(50, 600))
            pts = np.array([P(xs[j], 639 - 8.0 * (3 + j)) for j in range(n)], dtype=object)
            lanes.append(L(np.float32(rng.uniform(0.02, 1)), 3, 3 + n, 16.0, 16.0, 1, pts))
        return lanes
    rng = np.random.default_rng(0); a = r_conv(r_order(mk(RL, RP), 640), 3.0, 1.6875)
    rng = np.random.default_rng(0); b = hb.convert_lane_to_dict(hb.order_lane_x_axis(mk(hb.Lane, hb.Point), 640), 3.0, 1.6875)
    assert a == b


def test_regnet_plan_and_anchors():
    assert regnet_stage_plan(24, 36, 2.5, 30, 1, 8) == [(1, 24, 8), (1, 64, 8), (4, 152, 8), (10, 376, 8), (14, 936, 8)]
    assert [w for _, w, _ in regnet_stage_plan(24, 36, 2.5, 16, 1, 8)] == [24, 64, 152, 376]
    a = hb.make_anchors((640, 640), 2.0, [8, 16, 32, 64, 128], [2 ** 0.0, 2 ** 0.333, 2 ** 0.667], [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)])
    r = hydranet_ref.anchors((640, 640), 2.0, [3, 4, 5, 6, 7], [2 ** 0.0, 2 ** 0.333, 2 ** 0.667], [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)])
    assert a.shape == (1, 76725, 4) and np.array_equal(a, r.numpy())
    with pytest.raises(ValueError):
        hb.make_anchors((600, 640), 2.0, [8, 16, 32, 64, 128], [1.0], [(1.0, 1.0)])
    c = hb.LaneCodec(640, 640, 32, 80, True, 1, True)
    assert (c.feature_width, c.feature_height, c.points_per_anchor, c.interval, c.pt_nums_single_lane) == (20, 20, 4.0, 8.0, 162)


def test_tiling_heuristics():
    assert engine.choose_bn(24) == 32 and engine.choose_bn(112) == 112 and engine.choose_bn(936) == 256 and engine.choose_bn(512) == 256
    assert engine.choose_bn(1344) == 256 and engine.choose_bn(5) == 16
    for hw in ((20, 20), (40, 40), (160, 160), (5, 5), (12, 20)):
        th, tw = engine.choose_tile(*hw)
        assert th * tw == 128


def test_head_branches_are_contiguous_and_cover_the_heads():
    """Plan branches (hn_plan_set_branch): trunk ops 0, then seg 1, detection towers 2 and 3, lane 4 -- ascending and
    contiguous, which is what hn_plan_run requires to fork them after the trunk."""
    m = hb.HydraNet(big_cfg(128, 128)).eval()
    b = engine.Builder(m, 1, 128, 128, torch.device("cpu"), act_dtype=torch.float32).build(torch.zeros(1, 3, 128, 128))
    br = [getattr(op, "branch", 0) for op in b.ops]
    assert br == sorted(br) and set(br) == {0, 1, 2, 3, 4}
    for op in b.ops:
        want = {"seg": 1, "det.reg": 2, "det.cls": 3, "lane": 4}
        pre = next((k for k in want if op.name.startswith(k)), None)
        assert getattr(op, "branch", 0) == (want[pre] if pre else 0), op.name


def test_serving_mode_plan_key_and_cpu_refusals():
    """Host logic that needs no GPU: the plan cache key follows the fused decoders' configuration; CPU tensors are refused
    loudly by the pre-processing and the forward (no CPU fallback)."""
    m = hb.HydraNet(big_cfg(128, 128)).eval()
    assert m._fused_key() is None
    codec = hb.LaneCodec(128, 128, 32, 16, True, 1, True)
    m.fuse_postprocess(det=(0.4, 0.3), lane=(codec, 0.9, 80, False))
    k1 = m._fused_key()
    m.fuse_postprocess(det=(0.3, 0.3), lane=(codec, 0.9, 80, False))
    k2 = m._fused_key()
    m.fuse_postprocess(det=(0.3, 0.3))
    k3 = m._fused_key()
    assert k1 != k2 and k2 != k3 and k3[1] is None
    m.fuse_postprocess()
    assert m._fused_key() is None
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 128, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.preprocess(torch.zeros((1, 8, 8, 3), dtype=torch.uint8), (4, 4))
    with pytest.raises(TypeError):
        hb.preprocess(torch.zeros((1, 8, 8, 3)), (4, 4))
