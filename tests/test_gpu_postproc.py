"""GPU parity tests of the post-processing kernels (-m gpu): bit-exact against the numpy oracle
(oracle/postproc_ref.py, pinned against the live reference + torchvision) and against the golden
outputs of the live reference's own decoders."""
import os

import numpy as np
import pytest
import torch

import hydranet_b200 as hb
from hydranet_b200 import _native as nv
from hydranet_b200.heads import make_anchors
from oracle import postproc_ref as pr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ segmentation arg-max
@pytest.mark.parametrize("shape", [(2, 5, 640, 640), (1, 5, 1280, 1280), (3, 7, 33, 17), (1, 1, 8, 8), (0, 5, 16, 16)])
def test_seg_argmax_bit_exact(shape):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    if x.numel():
        x = torch.round(x * 4) / 4  # plenty of exact ties -> lowest index must win
    ref = torch.argmax(x, 1) if x.numel() else torch.zeros((0,) + shape[2:], dtype=torch.int64)
    out = hb.SegmentHeader.argmax(x.cuda())
    assert out.dtype == torch.int64 and torch.equal(out.cpu(), ref)


def test_seg_argmax_nan_semantics():
    x = torch.randn(1, 5, 16, 16)
    x[0, 3, 2, 2] = float("nan")
    x[0, 1, 5, 5] = float("nan"); x[0, 4, 5, 5] = float("nan")
    assert torch.equal(hb.SegmentHeader.argmax(x.cuda()).cpu(), torch.argmax(x, 1))


# ------------------------------------------------------------------ detection
def _det_compare(out, ref, ties_as_sets=False):
    assert len(out) == len(ref)
    for o, r in zip(out, ref):
        assert len(o["scores"]) == len(r["scores"])
        if len(r["scores"]) == 0:
            assert o["rois"].shape == (0,) and o["class_ids"].shape == (0,)
            continue
        assert o["rois"].dtype == np.float32 and o["class_ids"].dtype == np.int64 and o["scores"].dtype == np.float32
        assert np.array_equal(o["scores"], r["scores"])
        if not ties_as_sets:
            assert np.array_equal(o["class_ids"], r["class_ids"]) and np.array_equal(o["rois"], r["rois"])
        else:  # vanilla dispatch ends with an unstable score sort: equal-score runs compared as sets
            key = lambda d: sorted(map(tuple, np.concatenate([d["scores"][:, None], d["class_ids"][:, None].astype(np.float64), d["rois"]], 1).tolist()))
            assert key(o) == key(r)


@pytest.mark.parametrize("name,hw", [("big_128x128", (128, 128)), ("small_128x256", (128, 256))])
def test_det_decode_matches_live_reference_golden(name, hw):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    reg, cls, anc = (torch.from_numpy(g[k]).cuda() for k in ("regression", "classification", "anchors"))
    thr = float(g["det_thr"])
    boxes, scores, cids, count, _ = hb.DetectionHeader.decode_device(hw, reg, cls, anc, thr, 0.3, nms_mode=nv.NMS_AUTO_CPU)
    for i in range(2):
        k = int(count[i])
        assert k == len(g["det%d_scores" % i])
        assert np.array_equal(scores[i, :k].cpu().numpy(), g["det%d_scores" % i])
        assert np.array_equal(cids[i, :k].cpu().numpy(), g["det%d_class_ids" % i])
        # boxes: the reference's fp32 exp may differ from the correctly rounded one by 1 ulp
        assert np.allclose(boxes[i, :k].cpu().numpy(), g["det%d_rois" % i], rtol=3e-7, atol=1e-4)


def _synth_det(N, H, W, seed, hot=False):
    rng = np.random.default_rng(seed)
    anc = make_anchors((H, W), 2.0, [8, 16, 32, 64, 128], [2 ** 0.0, 2 ** 0.333, 2 ** 0.667], [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)])
    A = anc.shape[1]
    reg = rng.normal(0, 0.3, (N, A, 4)).astype(np.float32)
    if hot:  # random-init-like: every score within [0.47, 0.53] -> every anchor is a candidate
        cls = (0.5 + rng.uniform(-0.03, 0.03, (N, A, 9))).astype(np.float32)
    else:
        cls = (1.0 / (1.0 + np.exp(-rng.normal(0, 1, (N, A, 9))))).astype(np.float32)
    # force score ties so that the stable tie order is exercised
    n7 = cls[:, 1::7].shape[1]
    cls[:, 0:7 * n7:7] = cls[:, 1::7]
    return anc, reg, cls


@pytest.mark.parametrize("mode,pmode", [(nv.NMS_TRICK, "trick"), (nv.NMS_VANILLA, "vanilla"), (nv.NMS_AUTO_CUDA, None)])
def test_det_decode_nms_bit_exact_vs_oracle(mode, pmode):
    H = W = 256
    anc, reg, cls = _synth_det(3, H, W, 5)
    thr = 0.8
    out = []
    boxes, scores, cids, count, cand = hb.DetectionHeader.decode_device((H, W), torch.from_numpy(reg).cuda(), torch.from_numpy(cls).cuda(),
                                                                        torch.from_numpy(anc).cuda(), thr, 0.3, nms_mode=mode)
    for i in range(3):
        k = int(count[i])
        out.append({"rois": boxes[i, :k].cpu().numpy(), "class_ids": cids[i, :k].cpu().numpy(), "scores": scores[i, :k].cpu().numpy()}
                   if k else {"rois": np.array(()), "class_ids": np.array(()), "scores": np.array(())})
    ref = pr.det_postprocess(anc, reg, cls, H, W, thr, 0.3, device="cuda", mode=pmode)
    assert [int(c) for c in cand] == [int((cls[i].max(1) > np.float32(thr)).sum()) for i in range(3)]
    _det_compare(out, ref, ties_as_sets=(pmode == "vanilla"))


@pytest.mark.parametrize("force_seq", [0, 1])
def test_det_dense_duplicates_overflow_and_sequential_path(force_seq):
    """Hundreds of near-identical same-class boxes: more than 64 conflicting predecessors per box, so the
    parallel path flags the image and the sequential kernel takes over; force_seq=1 runs the sequential
    kernel for everything.  Both must equal the oracle bit for bit."""
    H = W = 256
    rng = np.random.default_rng(21)
    anc, reg, cls = _synth_det(2, H, W, 17)
    A = anc.shape[1]
    # image 0: 400 anchors moved onto (almost) the same box, same class, distinct scores
    idx = rng.choice(A, 400, replace=False)
    pre = pr.bbox_transform_clip(anc.reshape(-1, 4), reg, H, W)
    pre[0, idx] = np.array([60, 70, 150, 170], dtype=np.float32) + rng.uniform(-3, 3, (400, 4)).astype(np.float32)
    cls[0, idx] = 0.01
    cls[0, idx, 3] = rng.uniform(0.8, 0.99, 400).astype(np.float32)
    nv.lib.hn_det_force_sequential(force_seq)
    try:
        boxes, scores, cids, count, _ = hb.DetectionHeader.decode_device((H, W), None, torch.from_numpy(cls).cuda(), None, 0.75, 0.3,
                                                                         pre_boxes=torch.from_numpy(pre).cuda())
        torch.cuda.synchronize()
    finally:
        nv.lib.hn_det_force_sequential(0)
    ref = pr.det_postprocess(anc, reg, cls, H, W, 0.75, 0.3, device="cuda", pre_boxes=pre)
    out = [{"rois": boxes[i, :int(count[i])].cpu().numpy(), "class_ids": cids[i, :int(count[i])].cpu().numpy(),
            "scores": scores[i, :int(count[i])].cpu().numpy()} for i in range(2)]
    _det_compare(out, ref)


def test_det_iou_threshold_extremes():
    """thr = 0 (any overlap suppresses; pruning disabled -> sequential kernel) and thr = 0.9."""
    H = W = 128
    anc, reg, cls = _synth_det(1, H, W, 19)
    for thr in (0.0, 0.9):
        boxes, scores, cids, count, _ = hb.DetectionHeader.decode_device((H, W), torch.from_numpy(reg).cuda(), torch.from_numpy(cls).cuda(),
                                                                         torch.from_numpy(anc).cuda(), 0.7, thr)
        ref = pr.det_postprocess(anc, reg, cls, H, W, 0.7, thr, device="cuda")
        k = int(count[0])
        _det_compare([{"rois": boxes[0, :k].cpu().numpy(), "class_ids": cids[0, :k].cpu().numpy(), "scores": scores[0, :k].cpu().numpy()}], ref)


def test_det_public_decode_api_and_empty():
    H = W = 128
    anc, reg, cls = _synth_det(2, H, W, 7)
    imgs = torch.zeros(2, 3, H, W, device="cuda")
    cls[1] = 0.1  # image 1: nothing over threshold -> empty arrays like the reference
    out = hb.DetectionHeader.decode(imgs, torch.from_numpy(reg).cuda(), torch.from_numpy(cls).cuda(), torch.from_numpy(anc).cuda(), 0.7, 0.3)
    ref = pr.det_postprocess(anc, reg, cls, H, W, 0.7, 0.3, device="cuda")
    _det_compare(out, ref)
    assert out[1]["rois"].shape == (0,)
    assert hb.DetectionHeader.decode(None, None, None, None) is None


@pytest.mark.parametrize("size,n_anchor", [(640, 76725), (1280, 306900)])
def test_det_stress_max_anchors(size, n_anchor):
    """BASELINE config 5(i): every one of the 76 725 (640^2) / 306 900 (1280^2) anchors is an NMS candidate."""
    H = W = size
    anc, reg, cls = _synth_det(1, H, W, 11, hot=True)
    assert anc.shape[1] == n_anchor
    boxes, scores, cids, count, cand = hb.DetectionHeader.decode_device((H, W), torch.from_numpy(reg).cuda(), torch.from_numpy(cls).cuda(),
                                                                        torch.from_numpy(anc).cuda(), 0.3, 0.3)
    assert int(cand[0]) == n_anchor
    ref = pr.det_postprocess(anc, reg, cls, H, W, 0.3, 0.3, device="cuda")[0]  # more than 100 000 coordinates -> per-class
    k = int(count[0])
    out = {"rois": boxes[0, :k].cpu().numpy(), "class_ids": cids[0, :k].cpu().numpy(), "scores": scores[0, :k].cpu().numpy()}
    _det_compare([out], [ref], ties_as_sets=True)
    # size-independent properties: sorted by score, no surviving same-class pair above the IoU threshold
    assert np.all(np.diff(out["scores"]) <= 0)
    sub = np.random.default_rng(0).choice(k, size=min(k, 1500), replace=False)
    b, c = out["rois"][sub], out["class_ids"][sub]
    x1, y1 = np.maximum(b[:, None, 0], b[None, :, 0]), np.maximum(b[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(b[:, None, 2], b[None, :, 2]), np.minimum(b[:, None, 3], b[None, :, 3])
    inter = np.maximum(0, x2 - x1) * np.maximum(0, y2 - y1)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iou = inter / (area[:, None] + area[None, :] - inter)
    np.fill_diagonal(iou, 0)
    assert not np.any((iou > np.float32(0.3)) & (c[:, None] == c[None, :]))


def test_det_pre_boxes_identical_inputs():
    """NMS given identical pre-NMS boxes (north_star): feed the oracle's decoded boxes."""
    H = W = 256
    anc, reg, cls = _synth_det(2, H, W, 13)
    pre = pr.bbox_transform_clip(anc.reshape(-1, 4), reg, H, W)
    boxes, scores, cids, count, _ = hb.DetectionHeader.decode_device((H, W), None, torch.from_numpy(cls).cuda(), None, 0.75, 0.3,
                                                                     pre_boxes=torch.from_numpy(pre).cuda())
    ref = pr.det_postprocess(anc, reg, cls, H, W, 0.75, 0.3, device="cuda", pre_boxes=pre)
    out = [{"rois": boxes[i, :int(count[i])].cpu().numpy(), "class_ids": cids[i, :int(count[i])].cpu().numpy(),
            "scores": scores[i, :int(count[i])].cpu().numpy()} for i in range(2)]
    _det_compare(out, ref)


# ------------------------------------------------------------------ lanes
def _lane_compare(lanes, ref):
    assert len(lanes) == len(ref)
    for l, r in zip(lanes, ref):
        assert np.float32(l.prob) == np.float32(r["prob"]) and l.start_pos == r["start_pos"] and l.end_pos == r["end_pos"]
        assert l.ax == r["ax"] and l.ay == r["ay"]
        xs = np.array([p.x for p in l.lane], dtype=np.float32)
        ys = np.array([p.y for p in l.lane], dtype=np.float64)
        assert np.array_equal(xs, r["xs"]) and np.array_equal(ys, r["ys"])


@pytest.mark.parametrize("name,hw,stride", [("big_128x128", (128, 128), 32), ("small_128x256", (128, 256), 32)])
def test_lane_decode_matches_live_reference_golden(name, hw, stride):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    H, W = hw
    codec = hb.LaneCodec(W, H, stride, int(H / 8), True, 1, True)
    for b in range(2):
        prob = torch.from_numpy(g["lane%d_softmax" % b]).cuda()
        loc = torch.from_numpy(g["predict_loc"][b]).cuda()
        count, meta, p, xs, _ = hb.LaneHeader.decode_device(prob, loc, codec, 0.3, 30, False, cls_is_prob=True)
        lanes = hb.LaneHeader.lanes_from_device(count.cpu(), meta, p, xs, codec, 0)
        npts = g["lane%d_npts" % b]
        assert len(lanes) == len(npts)
        assert np.array_equal(np.array([l.prob for l in lanes], dtype=np.float32), g["lane%d_prob" % b])
        assert np.array_equal(np.array([l.start_pos for l in lanes]), g["lane%d_start" % b])
        assert np.array_equal(np.array([l.end_pos for l in lanes]), g["lane%d_end" % b])
        assert np.array_equal(np.concatenate([[p_.x for p_ in l.lane] for l in lanes]).astype(np.float32), g["lane%d_xs" % b])
        assert np.array_equal(np.concatenate([[p_.y for p_ in l.lane] for l in lanes]).astype(np.float64), g["lane%d_ys" % b])
        # logits path (softmax on the device): same lane set on this data
        lanes2 = hb.LaneHeader.decode(torch.from_numpy(g["predict_cls"][b]).cuda(), loc, codec, 0.3, 30, False)
        assert [l.start_pos for l in lanes2] == [l.start_pos for l in lanes]


@pytest.mark.parametrize("H,W,ppl,use_mean,thr,nms", [(640, 640, 80, False, 0.5, 80), (640, 640, 80, True, 0.3, 100),
                                                      (1280, 1280, 160, False, 0.5, 100), (128, 256, 16, False, 0.9, 80)])
def test_lane_decode_nms_bit_exact_vs_oracle(H, W, ppl, use_mean, thr, nms):
    """BASELINE config 5(ii): 400 and 1 600 anchors, cls logits N(0,3), loc N(0,2), end-pos U(0,ppl)."""
    rng = np.random.default_rng(H + ppl)
    codec = hb.LaneCodec(W, H, 32, ppl, True, 1, True)
    na = codec.feature_size
    N = 2
    logits = rng.normal(0, 3, (N, na, 2)).astype(np.float32)
    loc = rng.normal(0, 2, (N, na, 2 * ppl + 2)).astype(np.float32)
    loc[:, :, ppl] = rng.uniform(0, ppl, (N, na)).astype(np.float32)
    loc[:, :, ppl + 1] = rng.uniform(0, ppl, (N, na)).astype(np.float32)
    prob = pr.softmax2(logits)
    n5 = prob[:, 1::5].shape[1]
    prob[:, 0:5 * n5:5, 1] = prob[:, 1::5, 1]  # probability ties -> stable order
    count, meta, p, xs, cand = hb.LaneHeader.decode_device(torch.from_numpy(prob).cuda(), torch.from_numpy(loc).cuda(), codec, thr, nms,
                                                           use_mean, cls_is_prob=True)
    for b in range(N):
        ref = pr.lane_decode_nms(prob[b], loc[b], codec.feature_height, codec.feature_width, ppl, 32, codec.interval, W, H, thr, nms,
                                 use_mean, cls_is_prob=True)
        lanes = hb.LaneHeader.lanes_from_device(count.cpu(), meta, p, xs, codec, b)
        _lane_compare(lanes, ref)
    assert int(cand.max()) >= int(count.max())


def test_lane_empty_and_codec_decode_lane():
    codec = hb.LaneCodec(640, 640, 32, 80, True, 1, True)
    rng = np.random.default_rng(3)
    prob = np.zeros((400, 2), dtype=np.float32); prob[:, 0] = 1
    loc = rng.normal(0, 2, (400, 162)).astype(np.float32)
    assert hb.LaneHeader.decode(torch.from_numpy(prob).cuda() * 20, torch.from_numpy(loc).cuda(), codec, 0.5, 100) == []
    p2 = pr.softmax2(rng.normal(0, 3, (400, 2)).astype(np.float32))
    loc[:, 80] = 40; loc[:, 81] = 40
    cands = codec.decode_lane(torch.from_numpy(p2).cuda(), torch.from_numpy(loc).cuda(), 0.6)
    ref = pr.decode_lane(p2[:, 1], loc, 20, 20, 80, 32, 8.0, 640, 640, 0.6)
    _lane_compare(cands, ref)
    js = hb.LaneHeader.scale_to_org(cands[:3], 640, 640, 1920, 1080)
    assert "Lines" in js


# ------------------------------------------------------------------ host tails on the device
def test_invert_affine_device_matches_the_host_restatement():
    import hydranet_b200 as hb
    g = np.random.default_rng(3)
    N, A = 3, 500
    anchors = torch.from_numpy(g.uniform(0, 600, (A, 2)).astype(np.float32))
    anc = torch.cat([anchors, anchors + torch.from_numpy(g.uniform(8, 40, (A, 2)).astype(np.float32))], 1)[None].cuda()
    reg = torch.from_numpy(g.normal(0, 0.2, (N, A, 4)).astype(np.float32)).cuda()
    cls = torch.sigmoid(torch.from_numpy(g.normal(0, 2, (N, A, 9)).astype(np.float32))).cuda()
    x = torch.zeros(N, 3, 640, 640, device="cuda")
    metas = [(640, 640, 1920, 1080, 0, 0), (640, 640, 2560, 1440, 0, 0), (640, 640, 1570, 660, 0, 0)]
    plain = hb.DetectionHeader.decode(x, reg, cls, anc, 0.5, 0.3)
    want = hb.DetectionHeader.invert_affine(metas, hb.DetectionHeader.decode(x, reg, cls, anc, 0.5, 0.3))  # host restatement of detection.py:218-230 (in place)
    got = hb.DetectionHeader.decode(x, reg, cls, anc, 0.5, 0.3, metas=metas)
    assert any(len(p["rois"]) for p in plain)
    for a, b in zip(got, want):
        assert np.array_equal(a["rois"], b["rois"]) and np.array_equal(a["scores"], b["scores"]) and np.array_equal(a["class_ids"], b["class_ids"])
    got_f = hb.DetectionHeader.decode(x, reg, cls, anc, 0.5, 0.3, metas=0.3333)
    want_f = plain
    for p in want_f:  # `metas is float` is never true in the reference (an identity test against the type); the float branch restated directly
        if len(p["rois"]):
            p["rois"] = p["rois"] / np.float32(0.3333)
    for a, b in zip(got_f, want_f):
        assert np.array_equal(a["rois"], b["rois"])


@pytest.mark.parametrize("size,org", [((640, 640), (1920, 1080)), ((1280, 1280), (2560, 1440))])
def test_lane_scale_to_org_device_matches_the_host_path(size, org):
    import hydranet_b200 as hb
    W, H = size
    g = np.random.default_rng(11)
    codec = hb.LaneCodec(W, H, 32, int(H / 8), True, 1, True)
    fh, fw, ppl = codec.feature_height, codec.feature_width, codec.points_per_line
    cls = torch.from_numpy(g.normal(0, 3, (2, fh * fw, 2)).astype(np.float32)).cuda()
    loc = g.normal(0, 2, (2, fh * fw, 2 * ppl + 2)).astype(np.float32)
    loc[:, :, ppl] = g.uniform(0, ppl, (2, fh * fw))
    loc[:, :, ppl + 1] = g.uniform(0, ppl, (2, fh * fw))
    loc = torch.from_numpy(loc).cuda()
    count, meta, prob, xs, _ = hb.LaneHeader.decode_device(cls, loc, codec, 0.5, 60, False)
    assert int(count.min()) >= 2
    for i in range(2):
        lanes = hb.LaneHeader.lanes_from_device(count.cpu(), meta, prob, xs, codec, i)
        want = hb.LaneHeader.scale_to_org(lanes, W, H, org[0], org[1])  # host path (pinned against the live reference on the CPU)
        got = hb.LaneHeader.scale_to_org_device(count, meta, prob, xs, codec, W, H, org[0], org[1], index=i)
        assert len(got["Lines"]) == len(want["Lines"]) >= 2
        for a, b in zip(got["Lines"], want["Lines"]):
            assert a["score"] == b["score"] and type(a["score"]) is type(b["score"])
            assert len(a["points"]) == len(b["points"])
            for p, q in zip(a["points"], b["points"]):
                assert p["x"] == q["x"] and p["y"] == q["y"], (p, q)
                assert type(p["x"]) is type(q["x"]) and type(p["y"]) is type(q["y"])
