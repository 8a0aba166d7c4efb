"""ORACLE -- test infrastructure only.  Generates tests/golden/*.npz by RUNNING THE LIVE REFERENCE
(/root/reference, this container only) on the deterministic synthetic weights/inputs of
oracle/synth.py, and checks the oracle restatements against it while doing so.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hydranet_ref, postproc_ref, ref_live, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def cfg_variants():
    import yaml
    base = "/root/reference/model/cfgs/"
    big = yaml.safe_load(open(base + "hydranet_joint_big_backbone.yml"))
    small = yaml.safe_load(open(base + "hydranet_joint_small_backbone.yml"))
    out = {}
    for name, cfg, (h, w) in (("big_128x128", big, (128, 128)), ("small_128x256", small, (128, 256))):
        c = copy.deepcopy(cfg)
        c["dataloader"]["network_input_height"], c["dataloader"]["network_input_width"] = h, w
        out[name] = (c, h, w)
    return out


def lanes_to_arrays(lanes):
    return dict(prob=np.array([float(l.prob) for l in lanes], dtype=np.float32),
                start=np.array([l.start_pos for l in lanes], dtype=np.int32),
                end=np.array([l.end_pos for l in lanes], dtype=np.int32),
                ax=np.array([l.ax for l in lanes], dtype=np.float64),
                ay=np.array([l.ay for l in lanes], dtype=np.float64),
                npts=np.array([len(l.lane) for l in lanes], dtype=np.int32),
                xs=np.concatenate([np.array([p.x for p in l.lane], dtype=np.float32) for l in lanes]) if lanes else np.zeros(0, np.float32),
                ys=np.concatenate([np.array([p.y for p in l.lane], dtype=np.float64) for l in lanes]) if lanes else np.zeros(0, np.float64))


def main(out_dir=GOLD):
    ref_model, RefLaneCodec = ref_live.import_reference()
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(8)
    for name, (cfg, H, W) in cfg_variants().items():
        net = ref_model.HydraNet(cfg).eval()
        sd = synth.synth_state_dict(net.state_dict(), seed=1, seg_logit_gain=20.0)
        net.load_state_dict(sd)
        x = synth.synth_input(2, H, W, seed=3)
        with torch.no_grad():
            out = net(x)
            dep = net(x, mode="deploy")
            mine = hydranet_ref.forward(sd, cfg, x)
        # ---- pin the forward restatement: identical op sequence on the same CPU => bit-exact
        pairs = [("seg", out["seg"], mine["seg"]),
                 ("anchors", out["detection"]["anchors"], mine["detection"]["anchors"]),
                 ("regression", out["detection"]["regression"], mine["detection"]["regression"]),
                 ("classification", out["detection"]["classification"], mine["detection"]["classification"]),
                 ("predict_cls", out["lane"]["predict_cls"], mine["lane"]["predict_cls"]),
                 ("predict_loc", out["lane"]["predict_loc"], mine["lane"]["predict_loc"])]
        for k, a, b in pairs:
            assert a.shape == b.shape and torch.equal(a, b), "forward restatement differs from the live reference: " + k
        assert torch.equal(dep[0], torch.argmax(mine["seg"], 1))
        # ---- post-processing through the live reference
        det = out["detection"]
        cls_np = det["classification"].numpy()
        thr = float(np.quantile(cls_np.max(axis=2), 0.7))
        ref_det = ref_model.DetectionHeader.decode(x, det["regression"], det["classification"], det["anchors"], thr, 0.3)
        my_det = postproc_ref.det_postprocess(det["anchors"].numpy(), det["regression"].numpy(), cls_np, H, W, thr, 0.3, device="cpu")
        for r, m in zip(ref_det, my_det):
            assert np.array_equal(r["class_ids"], m["class_ids"]) and np.array_equal(r["scores"], m["scores"]), "det NMS restatement"
            assert np.allclose(r["rois"], m["rois"], rtol=3e-7, atol=1e-4), "det boxes restatement"
        lc = cfg["lane"]
        ppl = int(H / lc["interval"])
        codec = RefLaneCodec(W, H, lc["anchor_stride"], ppl, True, 1, True)
        lane_gold = {}
        for b in range(2):
            pc, pl = out["lane"]["predict_cls"][b], out["lane"]["predict_loc"][b]
            # random weights give P(lane) near 0.5: thresholds chosen so that candidates exist and NMS bites
            ref_lanes = ref_model.LaneHeader.decode(pc, pl, codec, 0.3, 30, False)
            prob = torch.softmax(pc, -1).numpy()
            mine_l = postproc_ref.lane_decode_nms(prob, pl.numpy(), codec.feature_height, codec.feature_width, ppl,
                                                  lc["anchor_stride"], codec.interval, W, H, 0.3, 30, False, cls_is_prob=True)
            ra = lanes_to_arrays(ref_lanes)
            assert len(ref_lanes) == len(mine_l), "lane count restatement %d vs %d" % (len(ref_lanes), len(mine_l))
            assert np.array_equal(ra["prob"], np.array([l["prob"] for l in mine_l], dtype=np.float32))
            assert np.array_equal(ra["start"], np.array([l["start_pos"] for l in mine_l], dtype=np.int32))
            assert np.array_equal(ra["xs"], np.concatenate([l["xs"] for l in mine_l]) if mine_l else np.zeros(0, np.float32))
            assert np.array_equal(ra["ys"], np.concatenate([l["ys"] for l in mine_l]) if mine_l else np.zeros(0))
            for k, v in ra.items():
                lane_gold["lane%d_%s" % (b, k)] = v
            lane_gold["lane%d_softmax" % b] = prob.astype(np.float32)
        gold = dict(seg=out["seg"].numpy(), seg_argmax=dep[0].numpy().astype(np.uint8), anchors=det["anchors"].numpy(),
                    regression=det["regression"].numpy(), classification=cls_np,
                    predict_cls=out["lane"]["predict_cls"].numpy(), predict_loc=out["lane"]["predict_loc"].numpy(),
                    det_thr=np.float64(thr))
        for i, r in enumerate(ref_det):
            gold["det%d_rois" % i], gold["det%d_class_ids" % i], gold["det%d_scores" % i] = r["rois"], r["class_ids"], r["scores"]
        gold.update(lane_gold)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **gold)
        print(name, "ok:", {k: v.shape for k, v in gold.items() if hasattr(v, "shape") and v.ndim > 0 and not k.startswith("lane")},
              "det kept", [len(r["scores"]) for r in ref_det], "lanes", [len(lane_gold["lane%d_prob" % b]) for b in range(2)])


def make_640_digest(out_dir=GOLD):
    """tests/golden/big_640x640_b1_digest.npz: the LIVE reference (CPU, IEEE fp32) at the default resolution, batch 1 -- the
    full-size pin for the GPU parity tests (the 128^2 fixtures above hold whole tensors; here the big tensors are stored as
    strided samples plus the complete arg-max map, < 1 MB).  Two weight sets: synthetic seed 1 (gain 20) and the reference's
    own random initialisation under torch.manual_seed(0)."""
    import yaml
    ref_model, _ = ref_live.import_reference()
    torch.set_num_threads(8)
    cfg = yaml.safe_load(open("/root/reference/model/cfgs/hydranet_joint_big_backbone.yml"))
    x = synth.synth_input(1, 640, 640, seed=5)
    blob = {}
    for tag in ("synth1", "init0"):
        torch.manual_seed(0)
        net = ref_model.HydraNet(cfg).eval()
        if tag == "synth1":
            net.load_state_dict(synth.synth_state_dict(net.state_dict(), seed=1, seg_logit_gain=20.0))
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        with torch.no_grad():
            out = net(x)
            mine = hydranet_ref.forward(sd, cfg, x)
        assert torch.equal(out["seg"], mine["seg"]) and torch.equal(out["detection"]["regression"], mine["detection"]["regression"])
        seg = out["seg"][0].numpy()
        top2 = np.sort(seg, axis=0)[-2:]
        blob[tag + ".seg_argmax"] = seg.argmax(0).astype(np.uint8)
        blob[tag + ".seg_gap_u8"] = np.clip((top2[1] - top2[0]) / (2e-2 * np.abs(seg).max()) * 64, 0, 255).astype(np.uint8)  # gap in 1/64 of the decisive bound
        blob[tag + ".seg_max"] = np.float32(np.abs(seg).max())
        for k, t, stride in (("seg", out["seg"], 101), ("regression", out["detection"]["regression"], 53), ("classification", out["detection"]["classification"], 53)):
            f = t.numpy().reshape(-1)
            blob["%s.%s.sample" % (tag, k)] = f[::stride].copy()
            blob["%s.%s.absmax" % (tag, k)] = np.float32(np.abs(f).max())
        blob[tag + ".predict_cls"] = out["lane"]["predict_cls"].numpy()
        blob[tag + ".predict_loc"] = out["lane"]["predict_loc"].numpy()
    np.savez_compressed(os.path.join(out_dir, "big_640x640_b1_digest.npz"), **blob)
    print("wrote big_640x640_b1_digest.npz", os.path.getsize(os.path.join(out_dir, "big_640x640_b1_digest.npz")) // 1024, "KB")


def reference_imagenet_normalize():
    """The reference's own ``imagenet_normalize`` (model/demo.py:26-40), compiled from its source text at generation
    time -- demo.py cannot be imported (argparse + checkpoint loading at module level)."""
    import ast
    src = open("/root/reference/model/demo.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "imagenet_normalize"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "demo.py", "exec"), ns)
    return ns["imagenet_normalize"]


def make_preprocess_golden(out_dir=GOLD):
    """tests/golden/preprocess.npz: demo.py:191-196 run LIVE (cv2 + the reference's normalise) on small frames."""
    import cv2
    from oracle import preprocess_ref
    norm = reference_imagenet_normalize()
    rng = np.random.default_rng(7)
    real = cv2.imread(sorted(__import__("glob").glob("/root/reference/model/demo/images/*.jpg"))[0])
    cases = {"down_45x80_to_48x32": (rng.integers(0, 256, (45, 80, 3), dtype=np.uint8), (48, 32)),
             "area_64x96_to_48x32": (rng.integers(0, 256, (64, 96, 3), dtype=np.uint8), (48, 32)),
             "up_37x53_to_96x64": (rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), (96, 64)),
             "real_crop_90x160_to_64x64": (np.ascontiguousarray(real[200:290, 300:460]), (64, 64))}
    blob = {}
    for name, (img, (w, h)) in cases.items():
        x = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        x = cv2.resize(x, (w, h))
        x = x.astype(np.float32)
        x = norm(x)
        x = np.transpose(x, (2, 0, 1))
        ref = torch.tensor(x).float().numpy()
        mine = preprocess_ref.preprocess(img, w, h)
        assert np.array_equal(ref, mine), name  # pins the restatement while generating
        blob[name + ".img"] = img
        blob[name + ".size"] = np.array([w, h], dtype=np.int32)
        blob[name + ".out"] = ref
    np.savez_compressed(os.path.join(out_dir, "preprocess.npz"), **blob)
    print("wrote preprocess.npz:", {k: v.shape for k, v in blob.items() if k.endswith(".out")}, "cv2", cv2.__version__)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "digest640":
        make_640_digest()
    elif len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        make_preprocess_golden()
    else:
        main()
        make_preprocess_golden()
