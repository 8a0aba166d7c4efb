"""Per-phase timing of selected conv launches using the kernel's globaltimer stamps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    names = sys.argv[1:] or ["backbone.s0.b0.c1", "backbone.s3.b5.c1", "backbone.s4.b5.c1", "backbone.s4.b5.c2", "neck.c1.conv3_up.pw",
                             "seg.d1.p00", "seg.d2", "seg.d3.p00", "seg.d5.p00", "seg.d6", "seg.d7.p00", "seg.out", "lane.hidden",
                             "det.cls.l0.0.pw"]
    dev = torch.device("cuda", 0)
    from hydranet_b200 import _native as nv
    dbg = torch.zeros(1024 * 16, dtype=torch.int64, device=dev)
    nv.lib.hn_conv_set_debug_buffer(dbg.data_ptr())
    hb, m, cfg = bench.build_model(dev)
    x = torch.randn(32, 3, 640, 640, device=dev)
    with torch.no_grad():
        m(x)
        m(x)
    torch.cuda.synchronize()
    plan = m.plan(32, 640, 640, dev)
    sp = torch.cuda.current_stream(dev).cuda_stream
    out = open(os.path.join(ROOT, "gpurun_out", "conv_probe.txt"), "w")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i, op in enumerate(plan.ops):
        if op.kind != "conv" or op.name not in names:
            continue
        for _ in range(3):
            plan.run_range(i, i + 1, sp)
        torch.cuda.synchronize()
        dbg.zero_()
        e0.record()
        plan.run_range(i, i + 1, sp)
        e1.record()
        torch.cuda.synchronize()
        d = dbg.view(-1, 16).cpu()
        d = d[d[:, 0] > 0]
        t0 = d[:, 0].min()
        rel = (d[:, :16] - t0).float() / 1e3  # us
        tiles = d[:, 8].float()
        f = lambda c: "%.1f/%.1f" % (rel[:, c].median().item(), rel[:, c].max().item())
        out.write("%-22s ev %.1f us | ctas %d tiles/cta %.1f | taps %d bn %d st %d | start %s setup %s prod1 %s full1 %s mma1 %s accfull1 %s [ld1 %s math1 %s bar1 %s store1 %s slab2 %s] epi1 %s end %s | GFLOP %.1f\n" % (
            op.name, e0.elapsed_time(e1) * 1e3, d.shape[0], tiles.mean().item(), len(op.taps), op.bn, op.stages,
            f(0), f(1), f(2), f(3), f(4), f(5), f(9), f(10), f(11), f(12), f(13), f(6), f(7), 2e-9 * op.macs))
        out.flush()
    out.close()
    print(open(os.path.join(ROOT, "gpurun_out", "conv_probe.txt")).read())


if __name__ == "__main__":
    main()
