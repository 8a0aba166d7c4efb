"""Test utility: run the native plan op by op next to the CPU emulator, feeding every native op the
emulator's (fp32 -> bf16) inputs, so each kernel launch is checked in isolation."""
import torch

import emulator
from hydranet_b200.engine import Builder, Plan


def _tensors_in(op):
    if op.kind == "conv":
        ts = [v.t for v in op.src]
        if op.res is not None:
            ts.append(op.res.t)
        return ts
    if op.kind == "node":
        return [v.t for v in op.ins]
    if op.kind == "dw_multi":
        return [v.t for v in op.ins]
    if op.kind == "pool":
        return [op.vin.t]
    if op.kind == "lanefuse":
        return [op.p3.t, op.p4.t, op.p5.t, op.p6.t]
    if op.kind in ("se_pool", "se_fused"):
        return [op.x.t]
    if op.kind == "se_scale":
        return [op.x.t, op.scale]
    if op.kind == "gconv_se":
        return [op.vin.t]
    if op.kind == "stem":
        return [op.x]
    raise KeyError(op.kind)


def _tensors_out(op):
    if op.kind == "conv":
        return [op.out_t] + ([op.out2] if op.out2 is not None else [])
    if op.kind == "node":
        return [op.out.t]
    if op.kind == "dw_multi":
        return [op.outs[0].t]
    if op.kind == "pool":
        return [op.vout.t]
    if op.kind == "lanefuse":
        return [op.out.t]
    if op.kind == "se_pool":
        return [op.mean] + ([op.fc["gate"]] if op.fc is not None else [])
    if op.kind == "se_scale":
        return [op.x.t]
    if op.kind in ("se_fused", "gconv_se"):
        return [op.mean, op.fc["gate"], op.x.t]
    if op.kind == "stem":
        return [op.out.t]
    raise KeyError(op.kind)


def lockstep(model_gpu, model_cpu, x, log=None):
    """Returns a list of (index, name, kind, max_abs_err, ref_max_abs, rel) per op."""
    dev = next(model_gpu.parameters()).device
    B, _, H, W = x.shape
    plan = Plan(model_gpu, B, H, W, dev)
    plan.x.copy_(x.to(dev))
    cpu = Builder(model_cpu, B, H, W, torch.device("cpu"), act_dtype=torch.float32).build(x.clone())
    assert len(cpu.ops) == len(plan.ops)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rows = []
    for i, (oc, og) in enumerate(zip(cpu.ops, plan.ops)):
        assert oc.kind == og.kind and oc.name == og.name
        for tc, tg in zip(_tensors_in(oc), _tensors_in(og)):
            tg.copy_(tc.to(tg.dtype))
        emulator.RUNNERS[oc.kind](oc)
        plan.run_range(i, i + 1, stream)
        torch.cuda.synchronize(dev)
        worst = (0.0, 0.0)
        for tc, tg in zip(_tensors_out(oc), _tensors_out(og)):
            a, b = tc.float(), tg.float().cpu()
            if tc.dtype == torch.uint8:
                err, ref = float((a != b).float().mean()), 1.0
            else:
                err, ref = float((a - b).abs().max()), float(a.abs().max())
            if err >= worst[0]:
                worst = (err, ref)
        rel = worst[0] / max(worst[1], 1e-6)
        rows.append((i, oc.name, oc.kind, worst[0], worst[1], rel))
        if log is not None:
            log.write("%4d %-28s %-8s err %.4e ref %.4e rel %.3e\n" % rows[-1])
            log.flush()
    return rows
