"""Test utility: a pure-torch CPU interpreter of the engine's operator specs.

It executes the *same* op list the native plan would (``engine.Builder.ops``) with the semantics the
CUDA kernels implement -- tap-table gathers with out-of-bounds zero fill, block-diagonal grouped
tiles, parity-strided outputs, halo mirroring, sub-pixel seg epilogue -- in fp32 on the CPU.  Running
it against the oracle forward validates all host-side logic (BN folding, weight packing, views, tap
tables) without a GPU.  It is test infrastructure, not a fallback: nothing in the package imports it.
"""
import torch
import torch.nn.functional as F

import hydranet_b200  # noqa: F401
from hydranet_b200 import _native as nv


def _act(x, act):
    if act == nv.ACT_RELU:
        return F.relu(x)
    if act == nv.ACT_SWISH:
        return x * torch.sigmoid(x)
    if act == nv.ACT_ELU:
        return F.elu(x)
    if act == nv.ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def _gather(v, ys, xs, c0, width=64):
    """[N, len(ys), len(xs), width] of view v at absolute coords, zero outside the view."""
    t = v.torch_view().float()
    N = v.N
    out = torch.zeros((N, len(ys), len(xs), width), dtype=torch.float32)
    yv = [(i, y) for i, y in enumerate(ys) if 0 <= y < v.H]
    xv = [(i, x) for i, x in enumerate(xs) if 0 <= x < v.W]
    if not yv or not xv or c0 >= v.C:
        return out
    cw = min(width, v.C - c0)
    if c0 < 0:
        return out
    yi = torch.tensor([i for i, _ in yv])
    ysrc = torch.tensor([y for _, y in yv])
    xi = torch.tensor([i for i, _ in xv])
    xsrc = torch.tensor([x for _, x in xv])
    blk = t[:, ysrc][:, :, xsrc][..., c0:c0 + cw]
    out[:, yi[:, None], xi[None, :], :cw] = blk
    return out


def run_conv(cs):
    w = cs.weight.float()
    bias = cs.bias.float()
    if cs.flat:
        v = cs.src[0]
        n_img, H, W = 1, 1, v.W
    else:
        n_img, H, W = cs.n_img, cs.out_h, cs.out_w
    ys, xs = list(range(H)), list(range(W))
    rows = w.shape[0]
    n_tiles = rows // cs.bn
    acc = torch.zeros((cs.src[0].N if not cs.flat else 1, H, W, rows), dtype=torch.float32)
    for k, (s, dy, dx, c0) in enumerate(cs.taps):
        wk = w[:, k * 64:(k + 1) * 64]
        if cs.grouped:
            for nt in range(n_tiles):
                a = _gather(cs.src[s], [y + dy for y in ys], [x + dx for x in xs], c0 + nt * cs.bn)
                acc[..., nt * cs.bn:(nt + 1) * cs.bn] += a @ wk[nt * cs.bn:(nt + 1) * cs.bn].t()
        else:
            a = _gather(cs.src[s], [y + dy for y in ys], [x + dx for x in xs], c0)
            acc += a @ wk.t()
    if cs.group_shift is not None:  # per-row-group affine instead of the plain bias (flat mode)
        M = W
        starts = [0] + list(cs.group_end[:-1])
        gidx = torch.zeros(M, dtype=torch.long)
        for g, (a0, a1) in enumerate(zip(starts, cs.group_end)):
            gidx[a0:a1] = g
        sc = cs.group_scale.float()[gidx] if cs.group_scale is not None else 1.0
        acc = (acc[0, 0] * sc + cs.group_shift.float()[gidx])[None, None]
    else:
        acc = acc + bias
    out_flat = cs.out_t.view(-1)
    if cs.epi == nv.EPI_SEGOUT:
        N = acc.shape[0]
        logits = cs.out_t
        for py in range(2):
            for px in range(2):
                r0 = (py * 2 + px) * 8
                logits[:, :, py::2, px::2] = acc[..., r0:r0 + cs.n_cls].permute(0, 3, 1, 2)
        if cs.out2 is not None:
            cs.out2.copy_(torch.argmax(logits, 1).to(torch.uint8))
        return
    acc = _act(acc[..., :cs.cout], cs.act)
    sn, sy, sx = cs.out_strides
    if cs.flat:
        M = W
        hw = cs.flat_hw
        m = torch.arange(M)
        n_i, pix = m // hw, m % hw
        off = cs.out_off + n_i * sn + pix * sx
        if cs.group_addr:
            starts = [0] + list(cs.group_end[:-1])
            off = torch.zeros(M, dtype=torch.long)
            for g, (a0, a1) in enumerate(zip(starts, cs.group_end)):
                ml = torch.arange(a1 - a0)
                off[a0:a1] = cs.out_off + (ml // cs.group_hw[g]) * sn + cs.group_out_base[g] + (ml % cs.group_hw[g]) * sx
        val = acc[0, 0]
        if cs.res is not None:
            r = cs.res
            roff = r.off + n_i * r.sn + pix * r.sx
            val = val + r.t.view(-1)[roff[:, None] + torch.arange(cs.cout)[None]].float()
            if cs.res_relu:
                val = F.relu(val)
        out_flat[off[:, None] + torch.arange(cs.cout)[None]] = val.to(cs.out_t.dtype)
        return
    N = acc.shape[0]
    OH, OW = H * cs.out_scale, W * cs.out_scale
    n_i = torch.arange(N)[:, None, None]
    Y = (torch.arange(H) * cs.out_scale + cs.out_oy)[None, :, None]
    X = (torch.arange(W) * cs.out_scale + cs.out_ox)[None, None, :]
    val = acc
    if cs.res is not None:
        r = cs.res
        roff = r.off + n_i * r.sn + Y * r.sy + X * r.sx
        val = val + r.t.view(-1)[roff[..., None] + torch.arange(cs.cout)].float()
        if cs.res_relu:
            val = F.relu(val)
    val = val.to(cs.out_t.dtype)
    cidx = torch.arange(cs.cout)

    def put(Yd, Xd, sel_y, sel_x):
        off = cs.out_off + n_i * sn + Yd * sy + Xd * sx
        out_flat[(off[..., None] + cidx)] = val[:, sel_y][:, :, sel_x]

    all_y, all_x = torch.arange(H), torch.arange(W)
    put(Y, X, all_y, all_x)
    if cs.halo != nv.HALO_NONE:
        Yl, Xl = (torch.arange(H) * cs.out_scale + cs.out_oy), (torch.arange(W) * cs.out_scale + cs.out_ox)

        def mirrors(coord, size):
            m = []  # (index into tile space, destination coordinate)
            for i, c in enumerate(coord.tolist()):
                if cs.halo == nv.HALO_REFLECT:
                    if c == 1:
                        m.append((i, -1))
                    if c == size - 2:
                        m.append((i, size))
                else:
                    if c == 0:
                        m.append((i, -1))
                    if c == size - 1:
                        m.append((i, size))
            return m
        my, mx = mirrors(Yl, OH), mirrors(Xl, OW)
        for (iy, dy_) in my:
            put(torch.tensor([dy_])[None, :, None], X, torch.tensor([iy]), all_x)
        for (ix, dx_) in mx:
            put(Y, torch.tensor([dx_])[None, None, :], all_y, torch.tensor([ix]))
        for (iy, dy_) in my:
            for (ix, dx_) in mx:
                put(torch.tensor([dy_])[None, :, None], torch.tensor([dx_])[None, None, :], torch.tensor([iy]), torch.tensor([ix]))


def _fetch(v, mode, H, W):
    t = v.torch_view().float().permute(0, 3, 1, 2)
    if mode == nv.IN_SAME:
        return t
    if mode == nv.IN_UP2:
        return F.interpolate(t, scale_factor=2, mode="nearest")
    return F.max_pool2d(F.pad(t, [0, 1, 0, 1]), 3, 2)


def run_node(ns):
    o = ns.out
    acc = None
    for i, v in enumerate(ns.ins):
        f = _fetch(v, ns.modes[i], o.H, o.W)
        acc = ns.ws[i] * f if acc is None else acc + ns.ws[i] * f
    if ns.swish:
        acc = acc * torch.sigmoid(acc)
    C = o.C
    w = ns.dw.t().reshape(C, 1, 3, 3)
    y = F.conv2d(F.pad(acc, [1, 1, 1, 1]), w, None, 1, 0, 1, C)
    o.torch_view().copy_(y.permute(0, 2, 3, 1))


def run_dw_multi(ds):
    for vin, vout in zip(ds.ins, ds.outs):
        C = vout.C
        t = vin.torch_view().float().permute(0, 3, 1, 2)
        w = ds.dw.t().reshape(C, 1, 3, 3)
        y = F.conv2d(F.pad(t, [1, 1, 1, 1]), w, None, 1, 0, 1, C)
        vout.torch_view().copy_(y.permute(0, 2, 3, 1))


def run_pool(ps):
    t = ps.vin.torch_view().float().permute(0, 3, 1, 2)
    y = F.max_pool2d(F.pad(t, [0, 1, 0, 1]), 3, 2) if ps.mode == nv.POOL_ZERO_RB else F.max_pool2d(t, 3, 2, 1)
    ps.vout.torch_view().copy_(y.permute(0, 2, 3, 1))


def run_lanefuse(ls):
    g = lambda v: v.torch_view().float().permute(0, 3, 1, 2)
    mp = lambda t: F.max_pool2d(t, 3, 2, 1)
    up = lambda t, f=2: F.interpolate(t, scale_factor=f, mode="nearest")
    if ls.stride == 32:
        y = torch.cat([mp(mp(g(ls.p3))), mp(g(ls.p4)), g(ls.p5), up(g(ls.p6))], 1)
    else:
        y = torch.cat([mp(g(ls.p3)), up(g(ls.p5)), g(ls.p4), up(g(ls.p6), 4)], 1)
    ls.out.torch_view().copy_(y.permute(0, 2, 3, 1))


def run_se_pool(ss):
    x = ss.x.torch_view()
    ss.mean.copy_(x.float().mean(dim=(1, 2)).to(ss.mean.dtype))
    if ss.fc is not None:  # FC1 + ReLU -> FC2 + sigmoid on the stored (rounded) mean; hidden rounded to the activation dtype
        f = ss.fc
        hid = F.relu(ss.mean.float() @ f["w1"].float().t() + f["b1"].float()).to(ss.mean.dtype)
        f["gate"].copy_(torch.sigmoid(hid.float() @ f["w2"].float().t() + f["b2"].float()).to(f["gate"].dtype))


def run_se_scale(ss):
    x = ss.x.torch_view()
    x.copy_((x.float() * ss.scale.float()[:, None, None, :]).to(x.dtype))


def run_se_fused(ss):
    run_se_pool(ss)
    x = ss.x.torch_view()
    x.copy_((x.float() * ss.fc["gate"].float()[:, None, None, :]).to(x.dtype))


def run_gconv_se(ss):
    xin = ss.vin.torch_view().float().permute(0, 3, 1, 2)
    y = F.relu(F.conv2d(xin, ss.w_ref.float().cpu(), ss.b_ref.float().cpu(), padding=1, groups=xin.shape[1] // 8))
    ss.x.torch_view().copy_(y.permute(0, 2, 3, 1))
    run_se_fused(ss)


def run_stem(st):
    w = st.w.reshape(3, 3, 3, 32).permute(3, 0, 1, 2)
    y = F.relu(F.conv2d(st.x, w, st.b, 2, 1))
    st.out.torch_view().copy_(y.permute(0, 2, 3, 1))


RUNNERS = {"conv": run_conv, "node": run_node, "dw_multi": run_dw_multi, "pool": run_pool, "lanefuse": run_lanefuse, "se_pool": run_se_pool, "se_scale": run_se_scale, "se_fused": run_se_fused, "gconv_se": run_gconv_se, "stem": run_stem}


def run_ops(ops):
    for op in ops:
        RUNNERS[op.kind](op)
