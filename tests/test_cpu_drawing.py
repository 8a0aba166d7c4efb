"""Host-side drawing helpers of the facade (LaneHeader.visual lanedetect.py:126-178, DetectionHeader.display
detection.py:247-252 -> display.py:53-84): CPU only."""
import numpy as np

import hydranet_b200 as hb


def test_visual_draws_and_filters():
    img = np.zeros((400, 600, 3), dtype=np.uint8)
    slanted = {"score": 0.93, "points": [{"x": 50.0 + 4 * i, "y": 380.0 - 3 * i} for i in range(60)]}
    vertical = {"score": 0.99, "points": [{"x": 300.0 + 0.25 * i, "y": 380.0 - 3 * i} for i in range(60)]}  # ~85 degrees
    short = {"score": 0.5, "points": [{"x": 10.0, "y": 10.0}]}
    out = hb.LaneHeader.visual([img.copy()], [[slanted, vertical, short]], 600)
    assert len(out) == 1 and out[0].shape == img.shape
    assert out[0][380 - 3 * 30, 50 + 4 * 30].tolist() == [255, 255, 0]     # on the slanted lane
    assert out[0][200, 315].tolist() == [0, 0, 0]                           # the vertical lane was filtered (> 65 degrees)
    kept = hb.LaneHeader.visual([img.copy()], [[vertical]], 600, filter_vertical=False)
    assert kept[0][200, 315].tolist() == [255, 255, 0]


def test_display_scales_boxes_and_keeps_empty_frames():
    frames = [np.zeros((200, 400, 3), dtype=np.uint8), np.zeros((200, 400, 3), dtype=np.uint8)]
    empty = {'rois': np.array(()), 'class_ids': np.array(()), 'scores': np.array(())}
    one = {'rois': np.array([[10.7, 20.2, 50.9, 60.1]], dtype=np.float32), 'class_ids': np.array([1]), 'scores': np.array([0.75], dtype=np.float32)}
    out = hb.DetectionHeader.display([empty, one], list(frames), ["car", "bus"], (400, 200), (100, 100))
    assert out[0] is frames[0] and not out[0].any()
    assert out[1] is not frames[1] and out[1].any()
    ys, xs = np.nonzero(out[1].any(axis=2))
    assert xs.max() <= 4 * 50 + 2 and ys.max() <= 2 * 60 + 2  # (int(50.9) * 400/100, int(60.1) * 200/100) + line width
    assert hb.DetectionHeader.display([empty, empty], list(frames), ["car"], (400, 200), (100, 100))[0] is frames[0]
