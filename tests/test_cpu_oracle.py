"""CPU suite (-m "not gpu"): the oracle against the golden vectors of the live reference, and -- where
/root/reference exists -- against the live reference itself."""
import os
import tempfile

import numpy as np
import pytest
import torch

from hydranet_b200.config import big_cfg, small_cfg
from hydranet_b200.model import HydraNet
from oracle import hydranet_ref, postproc_ref as pr, ref_live, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("big_128x128", big_cfg(128, 128), (128, 128)), ("small_128x256", small_cfg(256, 128), (128, 256))]


@pytest.mark.parametrize("name,cfg,hw", CASES)
def test_oracle_forward_matches_golden(name, cfg, hw):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    m = HydraNet(cfg).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=1, seg_logit_gain=20.0)
    x = synth.synth_input(2, hw[0], hw[1], seed=3)
    with torch.no_grad():
        out = hydranet_ref.forward(sd, cfg, x)
    for k, t in (("seg", out["seg"]), ("regression", out["detection"]["regression"]),
                 ("classification", out["detection"]["classification"]), ("predict_cls", out["lane"]["predict_cls"]),
                 ("predict_loc", out["lane"]["predict_loc"])):
        ref = g[k]
        # same op sequence; another host CPU may pick other oneDNN kernels, hence a (tight) tolerance
        assert np.allclose(t.numpy(), ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max()), k
    assert np.array_equal(out["detection"]["anchors"].numpy(), g["anchors"])
    assert (torch.argmax(out["seg"], 1).numpy() == g["seg_argmax"]).mean() > 0.9999


@pytest.mark.parametrize("name,cfg,hw", CASES)
def test_oracle_postproc_matches_golden(name, cfg, hw):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    H, W = hw
    det = pr.det_postprocess(g["anchors"], g["regression"], g["classification"], H, W, float(g["det_thr"]), 0.3, device="cpu")
    for i, d in enumerate(det):
        assert np.array_equal(d["scores"], g["det%d_scores" % i]) and np.array_equal(d["class_ids"], g["det%d_class_ids" % i])
        assert np.allclose(d["rois"], g["det%d_rois" % i], rtol=3e-7, atol=1e-4)
    ppl = H // 8
    for b in range(2):
        lanes = pr.lane_decode_nms(g["lane%d_softmax" % b], g["predict_loc"][b], H // 32, W // 32, ppl, 32, float(H) / ppl, W, H, 0.3, 30,
                                   False, cls_is_prob=True)
        assert np.array_equal(np.array([l["prob"] for l in lanes], dtype=np.float32), g["lane%d_prob" % b])
        assert np.array_equal(np.array([l["start_pos"] for l in lanes]), g["lane%d_start" % b])
        assert np.array_equal(np.array([l["end_pos"] for l in lanes]), g["lane%d_end" % b])
        assert np.array_equal(np.concatenate([l["xs"] for l in lanes]), g["lane%d_xs" % b])
        assert np.array_equal(np.concatenate([l["ys"] for l in lanes]), g["lane%d_ys" % b])
    assert np.array_equal(pr.seg_argmax(g["seg"]).astype(np.uint8), g["seg_argmax"])


def test_oracle_nms_matches_torchvision():
    """The third-party dependency on the path (torchvision 0.26, un-vendored): same keep order incl. ties."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(0)
    for n, dev_mode in ((300, "cpu"), (3000, "cpu"), (900, "cpu")):
        xy = rng.uniform(0, 200, (n, 2)).astype(np.float32)
        wh = rng.uniform(5, 80, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = np.round(rng.uniform(0, 1, n), 2).astype(np.float32)  # many ties
        idxs = rng.integers(0, 9, n)
        ref = tv.ops.boxes.batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(idxs), 0.3).numpy()
        got = pr.batched_nms(boxes, scores, idxs, 0.3, device=dev_mode)
        if boxes.size <= 4000:
            assert np.array_equal(ref, got)
        else:  # vanilla path ends in an unstable sort: compare as (score, index) multisets and score order
            assert np.array_equal(scores[ref], scores[got]) and set(ref.tolist()) == set(got.tolist())
        assert np.array_equal(tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy(), pr.nms(boxes, scores, 0.5))


@pytest.mark.skipif(not ref_live.available(), reason="live reference only exists in the build container")
def test_oracle_pinned_against_live_reference():
    """Re-runs oracle/make_golden.py: asserts the restatements equal the live reference bit for bit
    and that the committed fixtures are what the live reference produces today."""
    from oracle import make_golden
    with tempfile.TemporaryDirectory() as d:
        make_golden.main(d)
        for name, _, _ in CASES:
            a, b = np.load(os.path.join(d, name + ".npz")), np.load(os.path.join(GOLD, name + ".npz"))
            assert sorted(a.files) == sorted(b.files)
            for k in a.files:
                assert np.array_equal(a[k], b[k]), (name, k)


# ------------------------------------------------------------------ pre-processing
def _pre_cases():
    g = np.load(os.path.join(GOLD, "preprocess.npz"))
    names = sorted({k.rsplit(".", 1)[0] for k in g.files})
    return g, names


def test_preprocess_oracle_matches_golden():
    """oracle/preprocess_ref.py against tensors the LIVE pipeline (cv2 4.13 + the reference's imagenet_normalize) produced."""
    from oracle import preprocess_ref as pp
    g, names = _pre_cases()
    assert len(names) == 4
    for n in names:
        w, h = (int(v) for v in g[n + ".size"])
        assert np.array_equal(pp.preprocess(g[n + ".img"], w, h), g[n + ".out"]), n


def test_preprocess_oracle_matches_cv2_live():
    """the resize restatement against cv2 itself, bit for bit: down, up, exact 2x (INTER_AREA), identity, odd sizes."""
    cv2 = pytest.importorskip("cv2")
    from oracle import preprocess_ref as pp
    rng = np.random.default_rng(0)
    sizes = [(720, 1280, 640, 640), (1280, 1280, 640, 640), (360, 640, 640, 640), (717, 1283, 640, 640), (1080, 1920, 384, 640),
             (640, 640, 640, 640), (100, 37, 128, 256), (1280, 1920, 640, 640), (100, 100, 128, 100), (100, 100, 100, 128),
             (50, 50, 100, 100), (64, 64, 640, 640), (300, 300, 640, 640), (500, 700, 640, 640), (2, 2, 7, 5), (1, 9, 4, 4),
             (9, 1, 4, 4), (33, 65, 16, 32), (31, 63, 16, 32), (480, 854, 256, 512)]
    for (h, w, H, W) in sizes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(pp.resize_u8(img, W, H), cv2.resize(img, (W, H))), (h, w, H, W)
    lut = pp.normalize_lut()
    lv = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    assert np.array_equal(pp.imagenet_normalize(lv.astype(np.float32)).astype(np.float32).reshape(256, 3).T, lut)


# ------------------------------------------------------------------ training step: golden only
def test_train_step_golden_is_reproducible_and_product_refuses_cpu():
    """The pin of the training step: one step of train.py:241-269 run by the live reference (oracle/train_golden.py).
    Here: the fixture is well-formed, the live reference reproduces it when present, and the product refuses CPU tensors
    in train mode loudly instead of falling back (tests/test_gpu_train.py reproduces the fixture on the GPU)."""
    g = np.load(os.path.join(GOLD, "train_step_big_640x640.npz"))
    for k in ("loss_total", "loss_seg", "loss_det_cls", "loss_det_reg", "loss_lane_cls_pos", "loss_lane_cls_neg", "loss_lane_loc",
              "gradnorm.all", "gradnorm.backbone", "n_params_without_grad"):
        assert k in g.files and np.isfinite(g[k]), k
    assert int(g["n_params_without_grad"]) == 4  # neck.bifpn.0.p5_to_p6.* 
    total = 5.0 * g["loss_seg"] + (g["loss_det_cls"] + 50.0 * g["loss_det_reg"]) + (g["loss_lane_cls_pos"] + g["loss_lane_cls_neg"] + g["loss_lane_loc"])
    assert abs(total - g["loss_total"]) <= 1e-5 * abs(g["loss_total"])  # train.py:192-203 with the yml weights
    m = HydraNet(big_cfg(128, 128))
    m.train()
    with pytest.raises(RuntimeError):  # train mode runs on the native kernels too: no CPU fallback
        m(torch.zeros(1, 3, 128, 128))
    if not ref_live.available():
        return
    import copy
    import yaml
    from oracle import train_golden
    ref_model, _ = ref_live.import_reference()
    cfg = copy.deepcopy(yaml.safe_load(open("/root/reference/model/cfgs/hydranet_joint_big_backbone.yml")))
    torch.set_num_threads(8)
    saved = torch.Tensor.cuda
    try:
        if not torch.cuda.is_available():  # segmentation_loss.py:53 moves its class weights with Tensor.cuda()
            torch.Tensor.cuda = lambda self, *a, **k: self
        blob = train_golden.run_step(ref_model, cfg, 640, 640)
    finally:
        torch.Tensor.cuda = saved
    for k in g.files:
        a, b = np.asarray(blob[k], dtype=np.float64), np.asarray(g[k], dtype=np.float64)
        assert np.allclose(a, b, rtol=2e-3, atol=1e-6 * max(1.0, float(np.abs(b).max()))), k
