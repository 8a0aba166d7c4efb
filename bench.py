#!/usr/bin/env python
"""Benchmark of the HydraNet hot path (BASELINE.json metric: images/s forward + post-processing).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl native|reference]

A *step* is one pass of the hot path over one batch of synthetic images: native forward (backbone,
BiFPN, seg / detect / lane heads) + GPU post-processing (seg arg-max, box decode + NMS at the demo's
0.4 / 0.3, lane decode + NMS at 0.90 / 80).  Workload: BASELINE.json configs[1] -- big cfg, bf16,
batch 32 per GPU at 640x640, random-init weights (the reference ships none), synthetic images.
Multi-GPU: one process per GPU (torchrun), batch-sharded, no data-path collective (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/s fwd+postproc"
UNIT = "images/s"
DET_THR, LANE_THR = (0.4, 0.3), (0.90, 80)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of hn_conv_gemm_kernel per launch, MEASURED: the newest
    profiles/rNN_conv_traffic.json, written by tools/ncu_traffic.py from the committed ncu launch list of one step (ncu flushes the
    caches before every kernel, so this is an upper figure for the replayed step).  None when no profile has been taken."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d["dram_bytes_per_launch"], os.path.relpath(files[-1], ROOT)


def build_model(device):
    import hydranet_b200 as hb
    from hydranet_b200.config import big_cfg
    cfg = big_cfg()
    torch.manual_seed(0)
    m = hb.HydraNet(cfg).eval().to(device)
    return hb, m, cfg


def postproc(hb, m, out, codec, ws):
    if m._fused_post is not None:  # serving mode: the decoders ran inside forward, in the plan's detection / lane branches
        return m.postprocess_results()
    det = out["detection"]
    d = hb.DetectionHeader.decode_device((m.net_input_height, m.net_input_width), det["regression"], det["classification"],
                                         det["anchors"], DET_THR[0], DET_THR[1], workspace=ws)
    l = hb.LaneHeader.decode_device(out["lane"]["predict_cls"], out["lane"]["predict_loc"], codec, LANE_THR[0], LANE_THR[1], False)
    return d, l


def run_native(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hb, m, cfg = build_model(dev)
    m.use_graph = not args.no_graph
    m.static_outputs = True  # serving mode: zero-copy views of the plan's buffers (this loop snapshots what it downloads itself)
    from hydranet_b200 import _native as nv
    B, H, W = args.batch, 640, 640
    codec = hb.LaneCodec(W, H, cfg["lane"]["anchor_stride"], int(H / cfg["lane"]["interval"]), True, 1, True)
    if not args.no_fuse_postproc:
        m.fuse_postprocess(det=DET_THR, lane=(codec, LANE_THR[0], LANE_THR[1], False))
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]  # 157 MB each: larger than the 126 MB L2
    resident = [h.to(dev) for h in host]
    ws = torch.empty(nv.lib.hn_det_workspace_bytes(B, 76725), dtype=torch.uint8, device=dev)
    # a non-default stream: the forward is replayed as a CUDA graph (capture is illegal on the legacy stream)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    stream = torch.cuda.current_stream(dev)

    def step(x):
        with torch.no_grad():
            out = m(x)
            out["seg_cls_u8"] = m.seg_class_map()  # arg-max fused into the last seg conv's epilogue
            d, l = postproc(hb, m, out, codec, ws)
        return out, d, l

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput ----------------
    for i in range(args.warmup):
        step(resident[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(resident[i % 2])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- end to end through the public API with host buffers ----------------
    # Every step copies its own input batch host->device (pinned memory) and reads its results back; the
    # copies run on side streams so that step i+1's upload and step i-1's download overlap step i's compute.
    # Inputs travel as uint8 HWC camera frames (B x 640 x 640 x 3 = 39 MB instead of 157 MB of fp32 NCHW); the GPU pre-processing
    # kernel (demo.py:191-196: BGR->RGB, resize, normalise, HWC->CHW) writes straight into the plan's static input.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    host_u8 = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    xin = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    xin_f = [torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]  # pre-processed on the upload stream
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    pinned = [None, None]

    def upload(i):
        k = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[k])  # the compute that last read xin[k] has finished
            xin[k].copy_(host_u8[i % 2], non_blocking=True)
            hb.preprocess(xin[k], (W, H), out=xin_f[k])  # demo.py:191-196 on the GPU, on the upload stream: overlaps the previous step
            ev_in[k].record(s_in)

    segcopy = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
    pin_seg = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    pin_cnt = [torch.empty((2, B), dtype=torch.int32).pin_memory() for _ in range(2)]
    A_tot = 76725
    pin_det = [torch.empty((B, A_tot, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    pin_lane = [torch.empty((B, 400, 4 + 1 + 80), dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = [None, None]
    snaps = [None, None]
    ev_read = [torch.cuda.Event() for _ in range(2)]  # s_out has finished reading snapshot set k
    d2h_bytes = [0]

    def enqueue_small(k, seg_u8, d, l):
        """Stage 1 of the download (stream s_out): class map + per-image counts."""
        boxes, scores, cids, count, _ = d
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[k])
            for t in (seg_u8, count, boxes, scores, cids, l[0], l[1], l[2], l[3]):
                t.record_stream(s_out)  # per-call decoder outputs: keep the allocator from recycling them early
            pin_seg[k].copy_(seg_u8, non_blocking=True)
            pin_cnt[k][0].copy_(count, non_blocking=True)
            pin_cnt[k][1].copy_(l[0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_out)
        pending[k] = (ev, d, l)

    def finish(k):
        """Stage 2: once the counts are on the host, fetch exactly the kept detections / lanes (like the reference's
        `.cpu().numpy()` on the sliced tensors, detection_loss.py:99-103)."""
        if pending[k] is None:
            return
        ev, d, l = pending[k]
        pending[k] = None
        ev.synchronize()
        boxes, scores, cids, count, _ = d
        kmax, lmax = int(pin_cnt[k][0].max()), int(pin_cnt[k][1].max())
        with torch.cuda.stream(s_out):
            # pack on the device, then ONE contiguous pinned copy each (a strided D2H copy would go through a
            # pageable staging buffer and block the host)
            if kmax:
                pk = torch.empty((B, kmax, 6), dtype=torch.float32, device=dev)
                pk[:, :, 0:4] = boxes[:, :kmax]
                pk[:, :, 4] = scores[:, :kmax]
                pk[:, :, 5] = cids[:, :kmax]
                pin_det[k].view(-1)[:pk.numel()].copy_(pk.view(-1), non_blocking=True)
                pk.record_stream(s_out)
            if lmax:
                pl = torch.empty((B, lmax, 85), dtype=torch.float32, device=dev)
                pl[:, :, 0:4] = l[1][:, :lmax]
                pl[:, :, 4] = l[2][:, :lmax]
                pl[:, :, 5:] = l[3][:, :lmax]
                pin_lane[k].view(-1)[:pl.numel()].copy_(pl.view(-1), non_blocking=True)
                pl.record_stream(s_out)
            ev_read[k].record(s_out)
        d2h_bytes[0] = B * H * W + 2 * B * 4 + B * kmax * 24 + B * lmax * 85 * 4

    def e2e_loop(n, do_up=True, do_down=True):
        for k in range(2):
            ev_free[k].record(stream)
        if do_up:
            upload(0)
        for i in range(n):
            k = i % 2
            if do_up:
                if i + 1 < n:
                    upload(i + 1)
                stream.wait_event(ev_in[k])
            out, d, l = step(xin_f[k])
            ev_free[k].record(stream)
            if do_down:
                stream.wait_event(ev_read[k])  # the download that last used snapshot set k (two steps ago) is complete
                segcopy[k].copy_(out["seg_cls_u8"])  # the class map is the plan's static buffer: snapshot it for the download
                if m._fused_post is not None:
                    # serving mode: the decoders' outputs are static plan buffers as well, and the download of this step
                    # overlaps the next forward -- snapshot them on the compute stream (device-to-device, ~75 MB)
                    if snaps[k] is None:
                        snaps[k] = ([torch.empty_like(t) for t in d], [torch.empty_like(t) for t in l])
                    for dst, src in zip(snaps[k][0], d):
                        dst.copy_(src, non_blocking=True)
                    for dst, src in zip(snaps[k][1], l):
                        dst.copy_(src, non_blocking=True)
                    d, l = tuple(snaps[k][0]), tuple(snaps[k][1])
                ev_done[k].record(stream)
                enqueue_small(k, segcopy[k], d, l)
                finish(1 - k)  # while step i runs, complete the download of step i-1
        if do_down:
            finish((n - 1) % 2)
        s_out.synchronize()
        return None

    if args.e2e_probe and rank == 0:
        for name, up, down in (("compute only", False, False), ("upload+compute", True, False), ("compute+download", False, True),
                               ("full", True, True)):
            e2e_loop(4, up, down)
            torch.cuda.synchronize(dev)
            tp0 = time.perf_counter()
            e2e_loop(args.steps, up, down)
            torch.cuda.synchronize(dev)
            print("e2e probe %-18s %.3f ms/step" % (name, (time.perf_counter() - tp0) * 1e3 / args.steps), flush=True)

    e2e_loop(max(3, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    e2e_loop(args.steps)
    torch.cuda.synchronize(dev)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    t = torch.tensor([e2e_wall_ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_val = world * B * args.steps / (e2e_ms / 1e3)
    h2d = host_u8[0].numel()
    d2h = d2h_bytes[0]
    # host link speed, for context (pinned memory, 157 MB)
    tcp = []
    for _ in range(3):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        xin[0].copy_(host_u8[0], non_blocking=True)
        b_.record(stream)
        torch.cuda.synchronize(dev)
        tcp.append(a_.elapsed_time(b_))
    h2d_gbs = h2d / (min(tcp) * 1e-3) / 1e9

    # ---------------- roofline of the dominant kernel (tcgen05 implicit-GEMM conv) ----------------
    roof = cpu = None
    if rank == 0:
        plan = m.plan(B, H, W, dev)
        pk, pk_src = peaks()
        # per-launch device times: every op of the plan (both half-batch plans when the batch is split) between events
        units = [(part, i, op) for part in getattr(plan, "parts", [plan]) for i, op in enumerate(part.ops)]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(units) + 1)]
        sp = stream.cuda_stream
        tot = {}
        reps = 3
        for rep in range(reps + 1):
            plan.x.copy_(resident[rep % 2])
            # Park the GPU first (~15 ms spin on this stream) so that the host gets ahead and enqueues all launches and
            # events: otherwise the gap between two events is the host's per-launch cost (ctypes call + event record,
            # ~10 us), not the kernel, for every kernel shorter than that.
            torch.cuda._sleep(30_000_000)
            evs[0].record(stream)
            for j, (part, i, op) in enumerate(units):
                part.run_range(i, i + 1, sp)
                evs[j + 1].record(stream)
            torch.cuda.synchronize(dev)
            if rep == 0:
                continue
            for j, (part, i, op) in enumerate(units):
                k = (op.kind, op.group)
                a = tot.setdefault(k, [0.0, 0, 0])
                a[0] += evs[j].elapsed_time(evs[j + 1]) / reps
                a[1] += op.macs if rep == 1 else 0
                a[2] += 1 if rep == 1 else 0
        if args.dump_ops:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "op_times.txt"), "w") as f:
                for j, (part, i, op) in enumerate(units):
                    ms_i = evs[j].elapsed_time(evs[j + 1])
                    f.write("%4d %-30s %-8s %-9s %9.4f ms %8.3f GFLOP %8.2f TFLOP/s  (batch %d)\n" % (
                        j, op.name, op.kind, op.group, ms_i, 2e-9 * op.macs, 2e-9 * op.macs / max(ms_i, 1e-6), part.B))
        # HBM-bound kernels: algorithmic bytes (every input view read once + every output view written once) / event time
        def vbytes(v):
            return v.N * v.H * v.W * v.C * v.t.element_size()

        def op_bytes(op):
            if op.kind == "stem":
                return op.x.numel() * op.x.element_size() + vbytes(op.out)
            if op.kind == "node":
                return sum(vbytes(v) for v in op.ins) + vbytes(op.out)
            if op.kind == "dw_multi":
                return sum(vbytes(v) for v in op.ins) + sum(vbytes(v) for v in op.outs)
            if op.kind == "se_pool":
                return vbytes(op.x)
            if op.kind in ("se_scale", "se_fused", "gconv_se"):
                return 2 * vbytes(op.x)
            if op.kind == "lanefuse":
                return sum(vbytes(v) for v in (op.p3, op.p4, op.p5, op.p6, op.out))
            return 0
        hbm = {}
        for j, (part, i, op) in enumerate(units):
            b = op_bytes(op)
            if b:
                a = hbm.setdefault(op.kind, [0.0, 0.0])
                a[0] += b
                a[1] += evs[j].elapsed_time(evs[j + 1])
        hbm_peak = pk.get("hbm_gbs_sustained") or pk.get("hbm_gbs") or 0.0
        hbm = {k: {"GB/s": round(v[0] / (v[1] * 1e-3) / 1e9, 1), "ms": round(v[1], 3),
                   "frac": round(v[0] / (v[1] * 1e-3) / 1e9 / hbm_peak, 3) if hbm_peak else None} for k, v in sorted(hbm.items())}
        conv_ms = sum(v[0] for k, v in tot.items() if k[0] == "conv")
        conv_flops = 2.0 * sum(v[1] for k, v in tot.items() if k[0] == "conv")
        conv_n = sum(v[2] for k, v in tot.items() if k[0] == "conv")
        all_ms = sum(v[0] for v in tot.values())
        achieved = conv_flops / (conv_ms / 1e3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": "hn_conv_gemm_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
                "frac": round(achieved / peak, 4), "traffic": conv_traffic()[0], "traffic_source": conv_traffic()[1], "peak_source": pk_src + " (sustained: kernel timed inside a step)",
                "launches_per_step": conv_n, "ms_per_step_in_kernel": round(conv_ms, 3), "share_of_step": round(conv_ms / all_ms, 3),
                "algorithmic_gflop_per_launch_avg": round(conv_flops / conv_n / 1e9, 3),
                "breakdown_ms": {"%s/%s" % k: round(v[0], 3) for k, v in sorted(tot.items())},
                "hbm_bound_kernels": hbm, "hbm_peak_gbs": hbm_peak}
        # ---------------- CPU baseline: the oracle port of the reference path on the host cores ----------------
        cpu = cpu_baseline(m, cfg, seconds=args.cpu_seconds)

    lat = None
    if rank == 0 and not args.no_latency:
        # p50 batch-1 latency (BASELINE.json metric): forward (CUDA graph) + all three decoders, device time
        x1 = torch.randn(1, 3, H, W, device=dev)
        ws1 = torch.empty(nv.lib.hn_det_workspace_bytes(1, 76725), dtype=torch.uint8, device=dev)

        def step1():
            with torch.no_grad():
                out = m(x1)
                return postproc(hb, m, out, codec, ws1)
        for _ in range(5):
            step1()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(50):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step1()
            b_.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b_))
        ts.sort()
        # host cost of one forward call (Python facade + one graph launch), GPU kept busy so nothing waits on the device
        x1p = m.input_buffer(1, H, W, dev)
        torch.cuda.synchronize(dev)
        th0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(100):
                m(x1p)
        host_us = (time.perf_counter() - th0) * 1e4
        torch.cuda.synchronize(dev)
        lat = {"p50": round(ts[len(ts) // 2], 3), "p90": round(ts[int(len(ts) * 0.9)], 3), "unit": "ms", "batch": 1,
               "what": "forward (CUDA graph replay) + det/lane/seg post-processing, device time",
               "forward_host_us": round(host_us, 1)}

    if rank == 0:
        # own kernels per step: the plan (forward; + 12 detection + 1 lane kernels when the decoders are plan ops); the
        # detection NMS additionally launches 20 CUB kernels (two radix sorts, two scans), reported separately
        n_launch = m.plan(B, H, W, dev).n_launches + (0 if m._fused_post is not None else 12 + 1)
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "HydraNet big cfg bf16 inference, batch %d/GPU, 640x640, 3 heads + GPU post-processing "
                                       "(det 0.4/0.3, lane 0.90/80, seg argmax); random-init weights" % B,
                           "global_batch": world * B, "parallelism": "batch-sharded x%d, no collective" % world,
                           "l2_policy": "2 alternating input batches of 157 MB each (> 126 MB L2)"},
                "e2e": {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(e2e_ms / args.steps, 3), "h2d_link_gbs": round(h2d_gbs, 1),
                        "pipeline": "uint8 HWC frames up (39 MB) -> GPU pre-processing (upload stream) -> forward + decoders -> results down; "
                                    "upload of step i+1 and download of step i-1 overlap the compute of step i"},
                "gpu_launches": n_launch * args.steps, "library_launches": 20 * args.steps, "clocks": clocks, "latency_b1_ms": lat, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()



def run_train(args):
    """--mode train: BASELINE.json configs[3] -- one data-parallel training step per "step": train-mode forward (batch-stat
    BatchNorm), cal_loss (model.py:201-264), weighted total (train.py:192-203), backward (native dgrad / wgrad), bucketed NCCL
    gradient all-reduce overlapped with backward, Adam (train.py:147: lr 1e-5, wd 1e-8) in one kernel.  Weak scaling: `--batch`
    images per GPU.  Inputs and synthetic ground truth are resident on the device; e2e uploads them from pinned host memory every
    step and reads the loss back."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    import hydranet_b200 as hb
    from hydranet_b200.config import big_cfg
    from oracle import train_golden  # synthetic ground-truth generator only (config 4 formats); nothing of the oracle is timed
    cfg = big_cfg()
    torch.manual_seed(0)
    m = hb.HydraNet(cfg).to(dev).train()
    B, H, W = args.batch, 640, 640
    opt = hb.FusedAdam(m.parameters(), lr=cfg["train"]["lr"], weight_decay=cfg["train"]["weight_decay"])
    red = hb.GradAllReduce(m.parameters(), bucket_mb=25.0, model=m)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_x = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    gt_host = train_golden.synthetic_gt(B, H, W, H // 32, W // 32, H // 8, seed=5 + rank)
    gt_host = {k: v.pin_memory() for k, v in gt_host.items()}
    xs = [h.to(dev) for h in host_x]
    gt = {k: v.to(dev) for k, v in gt_host.items()}
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    # graph mode (default): forward + loss + backward (+ Adam when there is nothing to exchange) replayed as ONE CUDA graph,
    # gradients averaged after it by one coalesced NCCL all-reduce; --no-graph: eager launches with the bucketed exchange
    # overlapped with backward (GradAllReduce hooks)
    if args.no_graph:
        ts = hb.TrainStep(m, opt, graph=False, reducer=red)
    else:
        red.remove()
        ts = hb.TrainStep(m, opt, graph=True)

    def step(x, gtd, measure=False):
        return ts(x, gtd)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(xs[i % 2], gt)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        loss = step(xs[i % 2], gt)
    e1.record(stream)
    host_ms = (time.perf_counter() - t0) * 1e3  # host time to ENQUEUE the steps (Python + launches)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)
    # exposed (non-overlapped) part of the gradient exchange: time the compute stream waits for the comm stream at the end of backward
    exposed = None
    if world > 1:
        vals = []
        for i in range(3):
            if args.no_graph:  # time the compute stream spends waiting for the side-stream exchange after backward
                ts.reducer = None
                loss_, _ = ts._fwd_bwd(xs[i % 2], gt)
                red.finish(measure=True)
                ts.reducer = red
                opt.step()
                vals.append(red.exposed_ms)
            elif ts.reducer is not None:  # graph mode with the bucketed exchange captured inside the graph: not separable
                vals.append(None)
            else:  # graph mode: the exchange follows the graph, nothing overlaps it
                ts.static[0].copy_(xs[i % 2])
                ts.graph.replay()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(stream)
                ts._exchange()
                b_.record(stream)
                opt.step()
                b_.synchronize()
                vals.append(a_.elapsed_time(b_))
        exposed = sorted(vals)[1] if all(v is not None for v in vals) else None  # None: the exchange is inside the graph, not separable
    # e2e: pinned host -> device every step, loss read back
    barrier()
    xin = [torch.empty_like(xs[0]) for _ in range(2)]
    gin = [{k: torch.empty_like(v) for k, v in gt.items()} for _ in range(2)]
    s_in = torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        k = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[k])
            xin[k].copy_(host_x[k], non_blocking=True)
            for kk, v in gt_host.items():
                gin[k][kk].copy_(v, non_blocking=True)
            ev_in[k].record(s_in)

    def e2e_loop(n):
        for k in range(2):
            ev_free[k].record(stream)
        upload(0)
        losses_ = []
        for i in range(n):
            k = i % 2
            if i + 1 < n:
                upload(i + 1)
            stream.wait_event(ev_in[k])
            l = step(xin[k], gin[k])
            ev_free[k].record(stream)
            losses_.append(l.detach())
        return [float(v) for v in losses_]  # device -> host read of every step's loss

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    vals = e2e_loop(args.steps)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = host_x[0].numel() * 4 + sum(v.numel() * v.element_size() for v in gt_host.values())
    if rank == 0:
        n_param = sum(p.numel() for p in m.parameters())
        line = {"metric": "images/s train step (fwd+loss+bwd+allreduce+Adam)", "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "HydraNet big cfg training step, batch %d/GPU, 640x640, bf16 activations / fp32 master weights, 3 heads, "
                                       "cal_loss + Adam(lr 1e-5, wd 1e-8); random-init weights, synthetic ground truth (config 4)" % B,
                           "global_batch": world * B, "parallelism": "data-parallel x%d, NCCL gradient all-reduce (%.1f MB fp32), %s"
                                                                     % (world, n_param * 4 / 1e6, "bucketed, overlapped with backward (eager)" if args.no_graph
                                                                        else ("bucketed, captured inside the CUDA graph and overlapped with backward" if ts.reducer is not None
                                                                              else "one coalesced call after the CUDA-graph step")),
                           "launch": "eager" if args.no_graph else "CUDA graph (forward + loss + backward%s)" % (" + Adam" if world == 1 else ""),
                           "l2_policy": "2 alternating input batches (%d MB each)" % (host_x[0].numel() * 4 // 1000000)},
                "e2e": {"value": round(world * B * args.steps / (e2e_ms / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": round(e2e_ms / args.steps, 3)},
                "host_enqueue_ms_per_step": round(host_ms / args.steps, 3), "allreduce_exposed_ms": exposed, "final_loss": vals[-1], "clocks": clocks}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        ts.close()  # the graph may hold captured NCCL kernels: release it before the communicator goes away
        torch.cuda.synchronize(dev)
        dist.destroy_process_group()


def cpu_baseline(m, cfg, seconds=15.0, steps=None, batch=2):
    from oracle import cpu_baseline as cb
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    cb.run_once(sd, cfg, torch.randn(1, 3, 640, 640))  # warm-up
    t0 = time.perf_counter()
    n = 0
    x = torch.randn(batch, 3, 640, 640, generator=torch.Generator().manual_seed(0))
    while True:
        cb.run_once(sd, cfg, x)
        n += batch
        if (steps is not None and n >= steps * batch) or (steps is None and time.perf_counter() - t0 > seconds):
            break
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "%d images (batches of %d) in %.1f s: oracle/cpu_baseline.py = PyTorch fp32 CPU forward + torchvision "
                      "batched_nms + lane decode + argmax, all host threads" % (n, batch, dt)}


def _reference_cfg():
    """The keys oracle/hydranet_ref.forward reads, with the values of model/cfgs/hydranet_joint_big_backbone.yml."""
    return {"train": {"train_detect": True, "train_seg": True, "train_lane": True},
            "dataloader": {"network_input_width": 640, "network_input_height": 640},
            "backbone": {"group_width": 8, "stride": 2},
            "detection": {"num_classes": 9, "aspect_ratios_factor": [1.4, 0.7], "scales_factor": [0.0, 0.333, 0.667], "box_class_repeats": 3,
                          "pyramid_levels": 5, "anchor_scale": 2.0},
            "lane": {"anchor_stride": 32, "interval": 8, "num_classes": 2}}


def _reference_init(template, seed=0):
    """Random initialisation as the reference builds it: backbone convolutions N(0, sqrt(2 / fan_out)) (anynet.py:124-134), every
    other convolution PyTorch's default (kaiming_uniform(a=sqrt(5)) weights, U(+-1/sqrt(fan_in)) biases), BatchNorm gamma 1 / beta 0,
    running statistics 0 / 1, BiFPN fusion weights 1 (bifpn.py:104-123)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, t in template.items():
        shape = tuple(t.shape)
        stem = k.rsplit(".", 1)[0]
        if k.endswith("num_batches_tracked"):
            v = torch.zeros(shape, dtype=torch.int64)
        elif k.endswith("running_mean"):
            v = torch.zeros(shape)
        elif k.endswith("running_var") or "_w1" in k or "_w2" in k:
            v = torch.ones(shape)
        elif (stem + ".running_mean") in template:
            v = torch.ones(shape) if k.endswith(".weight") else torch.zeros(shape)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            if k.startswith("backbone."):
                v = torch.randn(shape, generator=g) * (2.0 / (shape[0] * shape[2] * shape[3])) ** 0.5
            else:
                v = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        else:  # conv bias
            w = template.get(stem + ".weight")
            fan_in = (w.shape[1] * w.shape[2] * w.shape[3]) if w is not None and w.dim() == 4 else shape[0]
            v = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        sd[k] = v
    return sd


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the live
    reference tree does not exist on the GPU box), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    cfg = _reference_cfg()
    # The state_dict is rebuilt from the committed key / shape list of the reference's big cfg (tests/golden/state_dict_keys_big.txt,
    # written next to the live reference): the reference arm never imports the native package, so none of its code or its .so is
    # mapped into this process.  Values: the reference's random initialisation (below), as in the native arm: detection scores
    # sit near 0.5, so every anchor passes the 0.4 threshold and the NMS sees the same maximum-candidate workload.
    template = {}
    for line in open(os.path.join(ROOT, "tests", "golden", "state_dict_keys_big.txt")):
        name, rest = line.split(" ", 1)
        shape = tuple(int(v) for v in rest[rest.index("(") + 1:rest.index(")")].replace(",", " ").split())
        template[name] = torch.empty(shape, dtype=torch.int64 if "int64" in rest else torch.float32)
    sd = _reference_init(template)
    torch.set_num_threads(os.cpu_count() or 1)
    batch = 2
    x = torch.randn(batch, 3, 640, 640, generator=torch.Generator().manual_seed(0))
    for _ in range(max(1, min(args.warmup, 3))):
        cb.run_once(sd, cfg, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cb.run_once(sd, cfg, x)
    dt = time.perf_counter() - t0
    v = batch * args.steps / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "HydraNet big cfg fp32 CPU (reference path), %d images per step of the batch-32 workload, 640x640, "
                                   "3 heads + post-processing (det 0.4/0.3, lane 0.90/80, seg argmax); random-init weights" % batch},
            "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d images per step x %d steps" % (batch, args.steps)},
            "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 32 for inference, 16 for training)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"], help="infer: configs[1]/[2] (the headline metric); train: configs[3], the training step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--dump-ops", action="store_true", help="write per-op device times to gpurun_out/op_times.txt")
    ap.add_argument("--e2e-probe", action="store_true", help="print the e2e loop time with upload / download switched off")
    ap.add_argument("--no-fuse-postproc", action="store_true", help="run the decoders after forward instead of inside the plan's head branches")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch-1 latency measurement")
    ap.add_argument("--no-graph", action="store_true", help="launch the forward kernel by kernel instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch is None:
        args.batch = 16 if args.mode == "train" else 32
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback (use --impl reference for the CPU path)")
    if args.mode == "train":
        return run_train(args)
    run_native(args)


if __name__ == "__main__":
    main()
