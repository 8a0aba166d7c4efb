"""The reference's demo loop body (model/demo.py:191-244) executed against the drop-in facade (-m gpu).

``_demo_body`` restates those lines call for call -- same pre-processing, same ``hydranet(img)``, same static
decode / scale_to_org / visual / decode / display calls with the demo's arguments -- so a maintainer's diff to run demo.py
on the B200 path is the import line only.  Weights are random-init (the reference ships none);
the frame is one of the reference's own demo images (tests/golden/demo_frame.jpg, 1570x660).
"""
import os

import cv2
import numpy as np
import pytest
import torch

import hydranet_b200 as hb
from hydranet_b200.config import big_cfg
from oracle import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

SEG_COLORS = {0: (0, 0, 0), 1: (128, 0, 128), 2: (255, 255, 255), 3: (0, 255, 255), 4: (0, 255, 0)}  # demo.py:91-96
OBJ_LIST = ["car", "truck", "bus", "person", "rider", "bike", "motor", "traffic_light", "traffic_sign"]


def _imagenet_normalize(img):  # demo.py:26-40
    return (img / np.array([255, 255, 255]) - np.array([0.485, 0.456, 0.406])) / np.array([0.229, 0.224, 0.225])


def _demo_body(hydranet, input_img, cfgs, lane_coder, lane_conf=0.90, lane_nms=80):
    net_input_width, net_input_height = cfgs["dataloader"]["network_input_width"], cfgs["dataloader"]["network_input_height"]
    net_input_size = (net_input_width, net_input_height)
    org_height, org_width = input_img.shape[0:2]
    org_size = (org_width, org_height)
    # demo.py:191-196
    img = cv2.cvtColor(input_img, cv2.COLOR_BGR2RGB)
    img = cv2.resize(img, net_input_size)
    img = _imagenet_normalize(img.astype(np.float32))
    img = np.expand_dims(np.transpose(img, (2, 0, 1)), axis=0)
    img = torch.tensor(img).cuda().float()
    # demo.py:202
    outputs = hydranet(img)
    imgs = [input_img]
    # demo.py:210-230
    cls_preds, loc_preds = outputs["lane"]['predict_cls'], outputs["lane"]['predict_loc']
    predict_jsons = []
    for batch_idx in range(len(imgs)):
        lane_nms_set = hydranet.laneheader.decode(cls_preds[batch_idx], loc_preds[batch_idx], lane_coder, lane_conf, lane_nms, False)
        predict_jsons.append(hydranet.laneheader.scale_to_org(lane_nms_set, net_input_width, net_input_height,
                                                              org_size[0], org_size[1])["Lines"])
    imgs = hydranet.laneheader.visual(imgs, predict_jsons, org_size[0], filter_vertical=True)
    # demo.py:232-235
    imgs = hydranet.segheader.decode(imgs, outputs["seg"], org_size, SEG_COLORS)
    # demo.py:238-244
    det = outputs["detection"]
    preds = hydranet.detectheader.decode(img, det["regression"], det["classification"], det["anchors"], conf_thres=0.4, iou_thres=0.3)
    imgs = hydranet.detectheader.display(preds, imgs, OBJ_LIST, org_size, (net_input_width, net_input_height))
    return imgs, predict_jsons, preds, outputs


def _setup(synthetic):
    cfgs = big_cfg()
    torch.manual_seed(0)
    hydranet = hb.HydraNet(cfgs=cfgs, onnx_export=False).cuda()
    if synthetic:
        hydranet.load_state_dict(synth.synth_state_dict(hydranet.state_dict(), seed=1, seg_logit_gain=20.0))
    hydranet.eval()
    lc = cfgs["lane"]
    coder = hb.LaneCodec(input_width=640, input_height=640, anchor_stride=lc["anchor_stride"], points_per_line=int(640 / lc["interval"]),
                         do_interpolate=lc["interpolate"], anchor_lane_num=lc["anchor_lane_num"], scale_invariance=lc["scale_invariance"])
    frame = cv2.imread(os.path.join(GOLD, "demo_frame.jpg"), cv2.IMREAD_UNCHANGED)
    assert frame is not None and frame.shape == (660, 1570, 3)
    return cfgs, hydranet, coder, frame


def test_demo_loop_body_random_init():
    """demo.py's thresholds on random-init weights: every anchor passes 0.4 (scores ~0.5), no lane passes 0.90."""
    cfgs, hydranet, coder, frame = _setup(False)
    with torch.no_grad():
        imgs, lanes, preds, outputs = _demo_body(hydranet, frame.copy(), cfgs, coder)
    assert len(imgs) == 1 and imgs[0].shape == frame.shape and imgs[0].dtype == np.uint8
    assert len(preds) == 1 and len(preds[0]["rois"]) > 1000 and preds[0]["rois"].shape[1] == 4
    assert lanes == [[]]
    assert not np.array_equal(imgs[0], frame)  # the seg overlay and the boxes were drawn
    # fresh tensors (reference semantics): a second forward must not overwrite what the caller holds
    keep = outputs["seg"].clone()
    with torch.no_grad():
        hydranet(torch.randn(1, 3, 640, 640, device="cuda"))
    assert torch.equal(keep, outputs["seg"])


def test_demo_loop_body_with_lanes_and_boxes():
    """Synthetic weights and a lane threshold low enough that lanes exist whatever the logits: visual() and display() both draw."""
    cfgs, hydranet, coder, frame = _setup(True)
    with torch.no_grad():  # make the lane head confident and its lanes long: foreground logit up, both end positions at 40 points
        hydranet.laneheader.conv_cls_conv[3].bias[1] += 8.0
        hydranet.laneheader.conv_up_conv[3].bias[80] = 40.0
        hydranet.laneheader.conv_down_conv[3].bias[80] = 40.0
    hydranet.refresh()  # parameters edited in place in eval mode
    with torch.no_grad():
        imgs, lanes, preds, _ = _demo_body(hydranet, frame.copy(), cfgs, coder, lane_conf=0.5)
    assert len(lanes[0]) >= 1 and all(set(l) == {"score", "points"} for l in lanes[0])
    assert imgs[0].shape == frame.shape
    os.makedirs(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out"), exist_ok=True)
    cv2.imwrite(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out", "demo_vis.jpg"), imgs[0])


def test_forward_on_non_current_device_or_default_stream():
    """model on one device while another is current; CUDA-graph replay from the legacy default stream."""
    cfgs = big_cfg(128, 128)
    torch.manual_seed(0)
    m = hb.HydraNet(cfgs).eval().cuda()
    x = torch.randn(1, 3, 128, 128, device="cuda")
    with torch.no_grad():
        a = m(x)["lane"]["predict_loc"]
        m.use_graph = True
        b = m(x)["lane"]["predict_loc"]  # current stream is the legacy default stream here
    assert torch.equal(a, b)
    if torch.cuda.device_count() > 1:
        m1 = hb.HydraNet(cfgs).eval()
        m1.load_state_dict(m.state_dict())
        m1 = m1.to("cuda:1")
        with torch.no_grad():
            c = m1(x.to("cuda:1"))["lane"]["predict_loc"]  # current device stays cuda:0
        assert torch.equal(a.cpu(), c.cpu())
