"""GPU parity of the pre-processing kernel (-m gpu): bit-exact against oracle/preprocess_ref.py (pinned against cv2 4.13
and the reference's imagenet_normalize) and against the golden tensors of the live pipeline (demo.py:191-196)."""
import os

import numpy as np
import pytest
import torch

import hydranet_b200 as hb
from oracle import preprocess_ref as pp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_preprocess_matches_live_reference_golden():
    g = np.load(os.path.join(GOLD, "preprocess.npz"))
    names = sorted({k.rsplit(".", 1)[0] for k in g.files})
    for n in names:
        w, h = (int(v) for v in g[n + ".size"])
        out = hb.preprocess(torch.from_numpy(g[n + ".img"]).cuda(), (w, h))
        assert out.shape == (1, 3, h, w) and out.dtype == torch.float32
        assert np.array_equal(out[0].cpu().numpy(), g[n + ".out"]), n


@pytest.mark.parametrize("h,w,H,W", [(720, 1280, 640, 640), (1280, 1280, 640, 640), (360, 640, 640, 640), (717, 1283, 640, 640),
                                     (1080, 1920, 384, 640), (640, 640, 640, 640), (100, 37, 128, 256), (2, 2, 7, 5),
                                     (1, 9, 4, 4), (9, 1, 4, 4), (31, 63, 16, 32)])
def test_preprocess_bit_exact_vs_oracle(h, w, H, W):
    """BASELINE config sizes (720p / 1080p camera frames -> 640x640, 384x640) and the edge cases the oracle is pinned on."""
    rng = np.random.default_rng(h * 7919 + w)
    imgs = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    out = hb.preprocess(torch.from_numpy(imgs).cuda(), (W, H)).cpu().numpy()
    for i in range(3):
        assert np.array_equal(out[i], pp.preprocess(imgs[i], W, H)), i


def test_preprocess_strided_rows_empty_batch_and_errors():
    rng = np.random.default_rng(5)
    big = torch.from_numpy(rng.integers(0, 256, (2, 60, 100, 3), dtype=np.uint8)).cuda()
    view = big[:, 5:50, 10:90]                      # rows keep the parent's pitch
    out = hb.preprocess(view, (32, 24)).cpu().numpy()
    for i in range(2):
        assert np.array_equal(out[i], pp.preprocess(view[i].cpu().numpy(), 32, 24))
    single = hb.preprocess(view[0], (32, 24))       # [h, w, 3] -> batch of one
    assert single.shape == (1, 3, 24, 32) and np.array_equal(single[0].cpu().numpy(), out[0])
    assert hb.preprocess(big[:0], (8, 8)).shape == (0, 3, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.preprocess(big.cpu(), (8, 8))
    with pytest.raises(TypeError):
        hb.preprocess(big.float(), (8, 8))


def test_preprocess_feeds_forward():
    """demo.py:191-202 end to end on device: frames -> preprocess -> HydraNet.forward."""
    from hydranet_b200.config import big_cfg
    from oracle import synth
    m = hb.HydraNet(big_cfg(128, 128)).eval().cuda()
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1, seg_logit_gain=20.0))
    rng = np.random.default_rng(9)
    frames = torch.from_numpy(rng.integers(0, 256, (2, 90, 160, 3), dtype=np.uint8)).cuda()
    x = hb.preprocess(frames, (m.net_input_width, m.net_input_height))
    with torch.no_grad():
        out = m(x)
    assert out["seg"].shape == (2, 5, 128, 128) and torch.isfinite(out["seg"]).all()
