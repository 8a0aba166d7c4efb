"""Phase times of the fused squeeze-excite launches (hn_se_fused_set_debug): python tools/se_fused_phases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hydranet_b200 import _native as nv
from hydranet_b200.engine import Buf

dev = torch.device("cuda")
names = ["staged", "conv", "pooled", "barrier1", "fc1+barrier2", "fc2", "scale"]
for (B, H, W, C, S) in [(32, 10, 10, 936, 240), (32, 20, 20, 376, 96)]:
    for conv in (False, True):
        bin_, bout = Buf(dev, torch.bfloat16, B, H, W, C), Buf(dev, torch.bfloat16, B, H, W, C)
        bin_.t.normal_(); bout.t.normal_()
        wq = torch.randn((C // 8, 10, 8, 8), device=dev).to(torch.bfloat16); bc = torch.zeros(C, device=dev)
        w1 = torch.randn((S, C), device=dev).to(torch.bfloat16) * 0.03; w2 = torch.randn((C, S), device=dev).to(torch.bfloat16) * 0.06
        b1 = torch.zeros(S, device=dev); b2 = torch.zeros(C, device=dev)
        partial = torch.zeros((B, 1, C), device=dev); counter = torch.zeros(B, dtype=torch.int32, device=dev)
        mean = torch.zeros((B, C), dtype=torch.bfloat16, device=dev); gate = torch.zeros_like(mean)
        se = nv.SePoolDesc(bout.interior().to_c(), 128, partial.data_ptr(), counter.data_ptr(), mean.data_ptr())
        se.S, se.w1, se.b1, se.w2, se.b2, se.gate = S, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gate.data_ptr()
        d = nv.GconvSeDesc(bin_.interior().to_c(), wq.data_ptr(), bc.data_ptr(), se)
        dbg = torch.zeros((B * 4, 8), dtype=torch.int64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        acc = torch.zeros(8); tot = 0.0
        for it in range(6):
            flush.zero_()
            bout.t.normal_()
            nv.lib.hn_se_fused_set_debug(dbg.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nv.check(nv.lib.hn_gconv_se_fwd(d, st) if conv else nv.lib.hn_se_fused_fwd(se, st))
            e1.record()
            torch.cuda.synchronize()
            nv.lib.hn_se_fused_set_debug(None)
            if it >= 2:
                t = dbg.cpu().double()
                start = t[:, 0].min()
                acc += torch.tensor([float((t[:, k] - start).max()) for k in range(8)]) / 1e3
                tot += e0.elapsed_time(e1) * 1e3
        acc /= 4; tot /= 4
        print("%dx%dx%d conv=%d: event %.1f us; last CTA reaches (us after first CTA start): %s; first-start spread %.1f us" % (
            H, W, C, conv, tot, ", ".join("%s %.1f" % (n, v) for n, v in zip(names, acc[1:].tolist())), float((t[:, 0].max() - t[:, 0].min()) / 1e3)))
