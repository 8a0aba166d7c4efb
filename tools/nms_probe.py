"""Phase breakdown of the detection NMS kernel on the bench workload (random-init scores)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

def main():
    dev = torch.device("cuda", 0)
    hb, m, cfg = bench.build_model(dev)
    from hydranet_b200 import _native as nv
    B = 32
    x = torch.randn(B, 3, 640, 640, device=dev)
    with torch.no_grad():
        out = m(x)
    det = out["detection"]
    cls = det["classification"]
    print("argmax class histogram of image 0:", torch.bincount(cls[0].argmax(1), minlength=9).tolist())
    dbg = torch.zeros(B * 16 * 8, dtype=torch.int64, device=dev)
    for it in range(3):  # clean timing first
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = hb.DetectionHeader.decode_device((640, 640), det["regression"], cls, det["anchors"], 0.4, 0.3)
        e1.record()
        torch.cuda.synchronize()
    print("decode+nms ms (no instrumentation)", e0.elapsed_time(e1))
    for it in range(1):
        dbg.zero_()
        nv.lib.hn_det_set_debug_buffer(dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = hb.DetectionHeader.decode_device((640, 640), det["regression"], cls, det["anchors"], 0.4, 0.3)
        e1.record()
        torch.cuda.synchronize()
    nv.lib.hn_det_set_debug_buffer(None)
    print("graph build: entries scanned %d, overlapping pairs %d (per image %.0f / %.0f)" % (
        int(dbg[0]), int(dbg[1]), float(dbg[0]) / B, float(dbg[1]) / B))
    d = dbg.view(B * 16, 8).cpu()
    d = d[d[:, 5] > 0]
    print("decode+nms ms", e0.elapsed_time(e1), "kept/img", r[3][:4].tolist(), "cand", r[4][:4].tolist())
    big = d[d[:, 5].argsort(descending=True)][:6]
    for row in big:
        print("seg size %6d kept %6d chunks %4d | cycles phase1 %9d phase2 %9d resolve %9d publish %9d" % (
            row[5], row[6], row[4], row[0], row[1], row[2], row[3]))

if __name__ == "__main__":
    main()
