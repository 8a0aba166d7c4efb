"""One bench step (forward + post-processing, batch 32) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Writes the op order (so kernel launches can be mapped back to
layers) to gpurun_out/op_order.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda", 0)
    hb, m, cfg = bench.build_model(dev)
    from hydranet_b200 import _native as nv
    codec = hb.LaneCodec(640, 640, 32, 80, True, 1, True)
    if os.environ.get("HN_FUSE_POST", "1") != "0":
        m.fuse_postprocess(det=bench.DET_THR, lane=(codec, bench.LANE_THR[0], bench.LANE_THR[1], False))
    x = torch.randn(B, 3, 640, 640, device=dev)
    ws = torch.empty(nv.lib.hn_det_workspace_bytes(B, 76725), dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(2):
            out = m(x)
            bench.postproc(hb, m, out, codec, ws)
        torch.cuda.synchronize()
        plan = m.plan(B, 640, 640, dev)
        plan = getattr(plan, 'parts', [plan])[0]  # op order of one half-batch plan when the batch is split
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "op_order.txt"), "w") as f:
            for i, op in enumerate(plan.ops):
                extra = ""
                if op.kind == "conv":
                    extra = "taps=%d bn=%d stages=%d cout=%d flat=%d tile=%s src=%d" % (len(op.taps), op.bn, op.stages, op.cout, op.flat, op.tile, len(op.src))
                f.write("%4d %-30s %-8s launches=%d %s\n" % (i, op.name, op.kind, op.launches, extra))
        convs = [op.name for op in plan.ops if op.kind == "conv"]
        with open(os.path.join(ROOT, "gpurun_out", "conv_index.txt"), "w") as f:
            for want in ("seg.d2", "seg.d3.p00", "seg.out", "backbone.s4.b5.c1"):
                f.write("%s %d\n" % (want, convs.index(want)))
        torch.cuda.profiler.start()
        out = m(x)
        bench.postproc(hb, m, out, codec, ws)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
