"""Every training-step operator (forward AND backward) against plain PyTorch fp32 autograd on the same bf16-rounded
inputs (-m gpu).  Tolerances: bf16 storage of activations / gradients -> max |err| <= 2e-2 of the tensor's max magnitude
(usually ~4e-3); fp32 statistics and parameter gradients of the HBM-bound ops <= 2e-3."""
import pytest
import torch
import torch.nn.functional as F

import hydranet_b200  # noqa: F401
from hydranet_b200 import _native as nv
from hydranet_b200 import losses
from hydranet_b200 import train as T

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def bf(t):
    return t.to(BF).float()


def nhwc(t):  # NCHW fp32 -> NHWC bf16
    return t.permute(0, 2, 3, 1).contiguous().to(BF)


def nchw(t):  # NHWC bf16 -> NCHW fp32
    return t.float().permute(0, 3, 1, 2)


def make_state(**convs):
    st = T.TrainState(None, torch.device(DEV, torch.cuda.current_device()))
    for name, (w, kind, kw) in convs.items():
        st.rec(name, w, kind, **kw)
    if convs:
        st._upload_table()
        st.pack()
    return st


def check(pairs, tol=2e-2):
    bad = {k: rel(a, b) for k, (a, b) in pairs.items() if not rel(a, b) <= tol}
    assert not bad, bad


@pytest.mark.parametrize("C,act,with_res,segs", [(24, nv.ACT_RELU, False, None), (152, nv.ACT_RELU, True, None), (112, nv.ACT_NONE, False, None),
                                                 (112, nv.ACT_SWISH, False, [700, 1000, 1100, 1130, 1150]), (936, nv.ACT_RELU, True, None)])
def test_batchnorm_train(C, act, with_res, segs):
    torch.manual_seed(1)
    st = make_state()
    R = segs[-1] if segs else 2 * 13 * 17
    z = (torch.randn(R, C, device=DEV) * 1.7 + 0.4).to(BF)
    res = torch.randn(R, C, device=DEV).to(BF) if with_res else None
    n = len(segs) if segs else 1
    bns = [torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.01).to(DEV) for _ in range(n)]
    for b in bns:
        b.weight.data.uniform_(0.5, 1.5)
        b.bias.data.normal_(0, 0.3)
        b.running_mean.normal_()
        b.running_var.uniform_(0.5, 2)
    ref_bns = [torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.01).to(DEV) for _ in range(n)]
    for a, b in zip(ref_bns, bns):
        a.load_state_dict(b.state_dict())
    zz = z.clone().requires_grad_()
    rr = res.clone().requires_grad_() if with_res else None
    y = T.BatchNormAct.apply(st, bns, act, segs, zz, rr, *([b.weight for b in bns] + [b.bias for b in bns]))
    gy = torch.randn_like(y)
    y.backward(gy)
    # reference
    zf = z.float().requires_grad_()
    rf = res.float().requires_grad_() if with_res else None
    outs, r0 = [], 0
    for i, e in enumerate(segs or [R]):
        u = ref_bns[i].train()(zf[r0:e].t().reshape(1, C, e - r0, 1)).reshape(C, e - r0).t()
        if with_res:
            u = u + rf[r0:e]
        u = F.relu(u) if act == nv.ACT_RELU else (u * torch.sigmoid(u) if act == nv.ACT_SWISH else u)
        outs.append(u)
        r0 = e
    yr = torch.cat(outs, 0)
    yr.backward(gy.float())
    pairs = {"y": (y, yr), "dz": (zz.grad, zf.grad)}
    if with_res:
        pairs["dres"] = (rr.grad, rf.grad)
    check(pairs)
    for a, b in zip(bns, ref_bns):
        check({"dgamma": (a.weight.grad, b.weight.grad), "dbeta": (a.bias.grad, b.bias.grad)}, 1e-2)
        check({"rmean": (a.running_mean, b.running_mean), "rvar": (a.running_var, b.running_var)}, 2e-3)


def _conv_case(kind, cin, cout, H, W, stride=1, bias=False, srcs=None, N=2):
    torch.manual_seed(2)
    k = 1 if kind.startswith("pw") else 3
    groups = cout // 8 if kind == "g3" else 1
    w = torch.nn.Parameter(torch.randn(cout, cin // groups, k, k, device=DEV) / (cin // groups * k * k) ** 0.5)
    b = torch.nn.Parameter(torch.randn(cout, device=DEV) * 0.1) if bias else None
    kw = {"stride": stride} if kind == "g3" else ({"src_channels": srcs} if srcs else {})
    st = make_state(c=(w, kind, kw))
    return st, st.recs["c"], w, b, groups


@pytest.mark.parametrize("cin,cout,bias,srcs", [(152, 376, False, None), (936, 936, False, None), (24, 24, True, None), (448, 448, False, [112] * 4),
                                                (112, 112, True, None), (32, 24, False, None)])
def test_conv1x1_fwd_dgrad_wgrad(cin, cout, bias, srcs):
    st, rec, w, b, _ = _conv_case("pw", cin, cout, 20, 20, bias=bias, srcs=srcs)
    x = torch.randn(2, 20, 20, cin, device=DEV).to(BF)
    xs = [t.clone().requires_grad_() for t in (x.split(srcs, dim=3) if srcs else [x])]
    y = T.Conv1x1.apply(st, rec, b, w, *xs)
    gy = torch.randn_like(y)
    y.backward(gy)
    xf = x.detach().float().requires_grad_()
    wf = bf(w.detach()).requires_grad_()
    bfp = b.detach().clone().requires_grad_() if bias else None
    yr = F.conv2d(xf.permute(0, 3, 1, 2), wf, bfp)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    check({"y": (nchw(y), yr), "dx": (torch.cat([t.grad for t in xs], 3), xf.grad), "dw": (w.grad, wf.grad)})
    if bias:
        check({"db": (b.grad, bfp.grad)}, 1e-2)


@pytest.mark.parametrize("cin,cout,H", [(64, 152, 40), (24, 64, 10)])
def test_conv1x1_stride2(cin, cout, H):
    st, rec, w, _, _ = _conv_case("pw_s2", cin, cout, H, H)
    x = torch.randn(2, H, H, cin, device=DEV).to(BF).requires_grad_()
    y = T.Conv1x1S2.apply(st, rec, w, x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xf = x.detach().float().requires_grad_()
    wf = bf(w.detach()).requires_grad_()
    yr = F.conv2d(xf.permute(0, 3, 1, 2), wf, None, 2)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    check({"y": (nchw(y), yr), "dx": (x.grad, xf.grad), "dw": (w.grad, wf.grad)})


@pytest.mark.parametrize("C,H,stride", [(24, 32, 1), (152, 20, 1), (376, 10, 1), (64, 40, 2), (152, 16, 2), (936, 10, 1)])
def test_grouped_conv3x3(C, H, stride):
    st, rec, w, _, groups = _conv_case("g3", C, C, H, H, stride=stride)
    x = torch.randn(2, H, H, C, device=DEV).to(BF).requires_grad_()
    y = T.GroupedConv3x3.apply(st, rec, w, x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xf = x.detach().float().requires_grad_()
    wf = bf(w.detach()).requires_grad_()
    yr = F.conv2d(xf.permute(0, 3, 1, 2), wf, None, stride, 1, 1, groups)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    check({"y": (nchw(y), yr), "dx": (nchw(x.grad), xf.grad.permute(0, 3, 1, 2)), "dw": (w.grad, wf.grad)})


@pytest.mark.parametrize("cin,cout,H,logits", [(112, 512, 20, False), (152, 128, 24, False), (64, 5, 32, True), (624, 512, 12, False)])
def test_conv3x3_padded_and_seggather(cin, cout, H, logits):
    st, rec, w, b, _ = _conv_case("c3", cin, cout, H, H, bias=True)
    cl = (cin // 2) // 8 * 8 if not logits else cin
    low = torch.randn(2, H // 2, H // 2, cl, device=DEV).to(BF).requires_grad_()
    skip = torch.randn(2, H, H, cin - cl, device=DEV).to(BF).requires_grad_() if cin > cl else None
    xp = T.SegGather.apply(low, skip)
    act = nv.ACT_NONE if logits else nv.ACT_ELU
    y = T.Conv3x3Padded.apply(st, rec, act, logits, w, b, xp)
    if logits:
        assert y.dtype == torch.float32
    gy = torch.randn_like(y)
    y.backward(gy)
    lf = low.detach().float().requires_grad_()
    sf = skip.detach().float().requires_grad_() if skip is not None else None
    wf = bf(w.detach()).requires_grad_()
    bfp = b.detach().clone().requires_grad_()
    cat = [F.interpolate(lf.permute(0, 3, 1, 2), scale_factor=2, mode="nearest")] + ([sf.permute(0, 3, 1, 2)] if sf is not None else [])
    u = F.conv2d(F.pad(torch.cat(cat, 1), [1, 1, 1, 1], mode="reflect"), wf, bfp)
    yr = u if logits else F.elu(u)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    pairs = {"y": (nchw(y) if not logits else y.permute(0, 3, 1, 2), yr), "dlow": (low.grad, lf.grad), "dw": (w.grad, wf.grad), "db": (b.grad, bfp.grad)}
    if skip is not None:
        pairs["dskip"] = (skip.grad, sf.grad)
    check(pairs)


def test_seggather_skip_only():
    x = torch.randn(2, 10, 14, 24, device=DEV).to(BF).requires_grad_()
    y = T.SegGather.apply(None, x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xf = x.detach().float().requires_grad_()
    yr = F.pad(xf.permute(0, 3, 1, 2), [1, 1, 1, 1], mode="reflect")
    yr.backward(gy.float().permute(0, 3, 1, 2))
    assert torch.equal(nchw(y), yr)
    check({"dx": (x.grad, xf.grad)}, 1e-2)


@pytest.mark.parametrize("stack", [False, True])
def test_depthwise3x3(stack):
    torch.manual_seed(3)
    st = make_state()
    C = 112
    w = torch.nn.Parameter(torch.randn(C, 1, 3, 3, device=DEV) / 3)
    shapes = [(2, 20, 20, C), (2, 10, 10, C), (2, 5, 5, C)] if stack else [(2, 13, 9, C)]
    xs = [torch.randn(s, device=DEV).to(BF).requires_grad_() for s in shapes]
    y = T.Depthwise3x3.apply(st, stack, w, *xs)
    gy = torch.randn_like(y)
    y.backward(gy)
    xfs = [x.detach().float().requires_grad_() for x in xs]
    wf = w.detach().clone().requires_grad_()
    outs = [F.conv2d(F.pad(xf.permute(0, 3, 1, 2), [1, 1, 1, 1]), wf, None, 1, 0, 1, C).permute(0, 2, 3, 1) for xf in xfs]
    yr = torch.cat([o.reshape(-1, C) for o in outs], 0) if stack else outs[0]
    yr.backward(gy.float())
    pairs = {"y": (y, yr), "dw": (w.grad, wf.grad)}
    for i, (a, b) in enumerate(zip(xs, xfs)):
        pairs["dx%d" % i] = (a.grad, b.grad)
    check(pairs)


@pytest.mark.parametrize("n_in", [2, 3])
def test_weighted_sum_swish(n_in):
    torch.manual_seed(4)
    st = make_state()
    ins = [torch.randn(2, 10, 10, 112, device=DEV).to(BF).requires_grad_() for _ in range(n_in)]
    p = torch.nn.Parameter(torch.tensor([1.0, 0.6, -0.2][:n_in], device=DEV))
    wn = torch.relu(p) / (torch.relu(p).sum() + 1e-4)
    a = T.WeightedSumSwish.apply(st, wn, *ins)
    ga = torch.randn_like(a)
    a.backward(ga)
    insf = [t.detach().float().requires_grad_() for t in ins]
    pf = p.detach().clone().requires_grad_()
    wnf = torch.relu(pf) / (torch.relu(pf).sum() + 1e-4)
    s = sum(wnf[k] * insf[k] for k in range(n_in))
    ar = s * torch.sigmoid(s)
    ar.backward(ga.float())
    pairs = {"a": (a, ar), "dp": (p.grad, pf.grad)}
    for k in range(n_in):
        if float(wnf[k]) != 0:
            pairs["din%d" % k] = (ins[k].grad, insf[k].grad)
        else:
            assert float(ins[k].grad.abs().max()) == 0
    check(pairs)


@pytest.mark.parametrize("mode,H,W", [(nv.RS_UP2, 5, 7), (nv.RS_POOL_ZERO, 20, 20), (nv.RS_POOL_ZERO, 10, 6), (nv.RS_POOL_NEGINF, 20, 20), (nv.RS_POOL_NEGINF, 9, 11)])
def test_resample(mode, H, W):
    torch.manual_seed(5)
    x = torch.randn(2, H, W, 112, device=DEV).to(BF).requires_grad_()
    y = T.Resample.apply(x, mode)
    gy = torch.randn_like(y)
    y.backward(gy)
    xf = x.detach().float().requires_grad_()
    xn = xf.permute(0, 3, 1, 2)
    if mode == nv.RS_UP2:
        yr = F.interpolate(xn, scale_factor=2, mode="nearest")
    elif mode == nv.RS_POOL_ZERO:
        yr = F.max_pool2d(F.pad(xn, [0, 1, 0, 1]), 3, 2)
    else:
        yr = F.max_pool2d(xn, 3, 2, 1)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    assert torch.equal(nchw(y), yr)
    check({"dx": (x.grad, xf.grad)}, 1e-2)


@pytest.mark.parametrize("C,S,H", [(24, 8, 16), (152, 38, 10), (936, 234, 5)])
def test_squeeze_excite(C, S, H):
    torch.manual_seed(6)
    st = make_state()
    w1 = torch.nn.Parameter(torch.randn(S, C, 1, 1, device=DEV) / C ** 0.5)
    b1 = torch.nn.Parameter(torch.randn(S, device=DEV) * 0.1)
    w2 = torch.nn.Parameter(torch.randn(C, S, 1, 1, device=DEV) / S ** 0.5)
    b2 = torch.nn.Parameter(torch.randn(C, device=DEV) * 0.1)
    x = (torch.randn(3, H, H, C, device=DEV) + 0.5).to(BF).requires_grad_()
    y = T.SqueezeExcite.apply(st, x, w1, b1, w2, b2)
    gy = torch.randn_like(y)
    y.backward(gy)
    xf = x.detach().float().requires_grad_()
    ps = [p.detach().clone().requires_grad_() for p in (w1, b1, w2, b2)]
    xn = xf.permute(0, 3, 1, 2)
    s = torch.sigmoid(F.conv2d(F.relu(F.conv2d(F.adaptive_avg_pool2d(xn, 1), ps[0], ps[1])), ps[2], ps[3]))
    yr = xn * s
    yr.backward(gy.float().permute(0, 3, 1, 2))
    check({"y": (nchw(y), yr), "dx": (x.grad, xf.grad)})
    check({"dw1": (w1.grad, ps[0].grad), "db1": (b1.grad, ps[1].grad), "dw2": (w2.grad, ps[2].grad), "db2": (b2.grad, ps[3].grad)}, 1e-2)


def test_head_conv_detection_layout_and_lane_layout():
    torch.manual_seed(7)
    B, C, na, k = 2, 112, 9, 9
    hws = [64, 16, 4]
    cout = na * k
    w = torch.nn.Parameter(torch.randn(cout, C, 1, 1, device=DEV) / C ** 0.5)
    b = torch.nn.Parameter(torch.randn(cout, device=DEV) * 0.1)
    st = make_state(h=(w, "pw", {}))
    rec = st.recs["h"]
    ends, acc, bases, a0 = [], 0, [], 0
    for hw in hws:
        acc += B * hw
        ends.append(acc)
        bases.append(a0 * k)
        a0 += hw * na
    total = sum(hws) * na
    x = torch.randn(acc, C, device=DEV).to(BF).requires_grad_()
    layout = ((B, total, k), 0, (total * k, cout), None, (ends, hws, bases))
    out = T.HeadConv.apply(st, rec, nv.ACT_SIGMOID, layout, None, w, b, x)
    g = torch.randn_like(out)
    out.backward(g)
    xf = x.detach().float().requires_grad_()
    wf, bfp = bf(w.detach()).requires_grad_(), b.detach().clone().requires_grad_()
    parts, r0 = [], 0
    for hw in hws:
        rows = xf[r0:r0 + B * hw].view(B, hw, C)
        parts.append(torch.sigmoid(rows @ wf.view(cout, C).t() + bfp).reshape(B, hw * na, k))
        r0 += B * hw
    ref = torch.cat(parts, 1)
    ref.backward(g)
    check({"out": (out, ref), "dx": (x.grad, xf.grad), "dw": (w.grad, wf.grad), "db": (b.grad, bfp.grad)})
    # lane layout: [B, fh*fw, cout], rows per image
    w2 = torch.nn.Parameter(torch.randn(81, 448, 1, 1, device=DEV) / 448 ** 0.5)
    b2 = torch.nn.Parameter(torch.randn(81, device=DEV) * 0.1)
    st2 = make_state(h=(w2, "pw", {}))
    x2 = torch.randn(B, 5, 5, 448, device=DEV).to(BF).requires_grad_()
    lay = ((B, 25, 81), 0, (25 * 81, 81), 25, None)
    o2 = T.HeadConv.apply(st2, st2.recs["h"], nv.ACT_NONE, lay, None, w2, b2, x2)
    g2 = torch.randn_like(o2)
    o2.backward(g2)
    x2f = x2.detach().float().requires_grad_()
    w2f, b2f = bf(w2.detach()).requires_grad_(), b2.detach().clone().requires_grad_()
    r2 = x2f.view(B, 25, 448) @ w2f.view(81, 448).t() + b2f
    r2.backward(g2)
    check({"out": (o2, r2), "dx": (x2.grad, x2f.grad), "dw": (w2.grad, w2f.grad), "db": (b2.grad, b2f.grad)})


def test_stem_conv():
    torch.manual_seed(8)
    st = make_state()
    w = torch.nn.Parameter(torch.randn(32, 3, 3, 3, device=DEV) / 27 ** 0.5)
    x = torch.randn(2, 3, 64, 96, device=DEV)
    z = T.StemConv.apply(st, x, w)
    gz = torch.randn_like(z)
    z.backward(gz)
    wf = w.detach().clone().requires_grad_()
    zr = F.conv2d(x, wf, None, 2, 1)
    zr.backward(gz.float().permute(0, 3, 1, 2))
    check({"z": (nchw(z), zr)})
    check({"dw": (w.grad, wf.grad)}, 2e-3)


def test_fused_adam_matches_torch():
    from hydranet_b200.optim import FusedAdam
    torch.manual_seed(9)
    shapes = [(936, 936, 1, 1), (32, 3, 3, 3), (112,), (5,), (3,), (512, 624, 3, 3)]
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a = FusedAdam(ps, lr=1e-3, weight_decay=1e-2)
    b = torch.optim.Adam(qs, lr=1e-3, weight_decay=1e-2)
    for step in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        assert float((p - q).abs().max()) <= 1e-6 * max(1.0, float(q.abs().max()))


def test_native_detection_loss_matches_the_pinned_torch_loss():
    """hn_det_loss (SURVEY f-3) against losses.detection_loss, which tests/test_cpu_losses.py pins against the live reference:
    loss values and the gradients w.r.t. classification and regression, incl. an image without boxes and ignored anchors."""
    import hydranet_b200 as hb
    from hydranet_b200 import losses
    from oracle import train_golden
    torch.manual_seed(0)
    B, A = 3, 76725
    gt = train_golden.synthetic_gt(B, 640, 640, 20, 20, 80, seed=5)
    gt["gt_det"][2, :, 4] = -1
    ann = gt["gt_det"].cuda()
    anc = torch.from_numpy(hb.make_anchors((640, 640), 2.0, [8, 16, 32, 64, 128], [2 ** 0, 2 ** 0.333, 2 ** 0.667], [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)])).cuda()
    cls0 = torch.sigmoid(torch.randn(B, A, 9, device=DEV) * 3)  # some scores outside the [1e-4, 1 - 1e-4] clamp
    reg0 = 0.3 * torch.randn(B, A, 4, device=DEV)
    res = {}
    for name in ("torch", "native"):
        cls, reg = cls0.clone().requires_grad_(), reg0.clone().requires_grad_()
        if name == "torch":
            c, r = losses.detection_loss(cls, reg, anc, ann)
            c, r = c.mean(), r.mean()
        else:
            c, r = losses.NativeDetectionLoss.apply(cls, reg, anc, ann)
        (2.0 * c + 50.0 * r).backward()
        res[name] = (float(c), float(r), cls.grad, reg.grad)
    a, b = res["native"], res["torch"]
    assert abs(a[0] - b[0]) <= 2e-6 * abs(b[0]) and abs(a[1] - b[1]) <= 2e-6 * abs(b[1]), (a[:2], b[:2])
    assert float((a[2] - b[2]).abs().max()) <= 2e-5 * float(b[2].abs().max())
    assert float((a[3] - b[3]).abs().max()) <= 2e-5 * float(b[3].abs().max())
    assert float(b[3].abs().max()) > 0 and float(b[2].abs().max()) > 0


@pytest.mark.parametrize("N,H,W,ratio,ignore_frac", [(3, 40, 56, 0.3, 0.1), (2, 64, 64, 0.3, 0.0), (2, 33, 17, 0.05, 0.85), (1, 640, 640, 0.3, 0.02)])
def test_native_seg_loss_matches_the_pinned_torch_loss(N, H, W, ratio, ignore_frac):
    """hn_seg_loss_fwd/bwd (cross-entropy + radix-select top-k) against losses.seg_loss, which is pinned
    against the live reference (tests/test_cpu_losses.py).  Continuous random logits: no ties at the threshold except the zeros
    of ignored pixels, which carry no gradient either way (third case: k exceeds the number of non-ignored pixels)."""
    g = torch.Generator().manual_seed(N * 100 + H)
    logits0 = (torch.randn((N, 5, H, W), generator=g) * 2.0).cuda()
    tgt = torch.randint(0, 5, (N, H, W), generator=g)
    tgt[torch.rand((N, H, W), generator=g) < ignore_frac] = 255
    tgt = tgt.cuda()
    w = torch.tensor([0.1, 0.5, 1.0, 5.0, 5.0], device="cuda")
    res = {}
    for name in ("torch", "native"):
        x = logits0.clone().requires_grad_()
        if name == "torch":
            loss = losses.seg_loss(x, tgt, w, True, ratio, False)
        else:
            loss = losses.NativeSegLoss.apply(x, tgt, w, ratio)
        (5.0 * loss).backward()
        res[name] = (float(loss), x.grad.clone())
    a, b = res["native"], res["torch"]
    assert abs(a[0] - b[0]) <= 2e-6 * abs(b[0]), (a[0], b[0])
    assert float(b[1].abs().max()) > 0
    assert float((a[1] - b[1]).abs().max()) <= 2e-5 * float(b[1].abs().max())
    assert int(((a[1] != 0) != (b[1] != 0)).sum()) == 0, "a different set of pixels was kept"


def test_native_seg_loss_ties_share_their_weight():
    """Identical pixels (exact ties at the threshold): the loss value equals torch's; the gradient is spread over the tied
    pixels instead of an arbitrary subset, and sums to the same total."""
    x0 = torch.zeros((1, 5, 8, 8), device="cuda")
    x0[:, 2] = 1.5  # every pixel identical
    tgt = torch.ones((1, 8, 8), dtype=torch.int64, device="cuda")
    w = torch.tensor([0.1, 0.5, 1.0, 5.0, 5.0], device="cuda")
    res = {}
    for name in ("torch", "native"):
        x = x0.clone().requires_grad_()
        loss = losses.seg_loss(x, tgt, w, True, 0.25, False) if name == "torch" else losses.NativeSegLoss.apply(x, tgt, w, 0.25)
        loss.backward()
        res[name] = (float(loss), x.grad.clone())
    assert abs(res["native"][0] - res["torch"][0]) <= 1e-6 * abs(res["torch"][0])
    gn, gt = res["native"][1], res["torch"][1]
    assert torch.allclose(gn.sum(dim=(2, 3)), gt.sum(dim=(2, 3)), rtol=1e-5, atol=1e-7)
    assert float(gn.abs().max()) <= float(gt.abs().max()) * 0.26  # 16 of 64 pixels kept: a quarter of the weight each


@pytest.mark.parametrize("B,na,L,n_pos", [(4, 400, 162, 5), (2, 100, 162, 0), (1, 64, 162, 40), (16, 400, 162, 3)])
def test_native_lane_loss_matches_the_pinned_torch_losses(B, na, L, n_pos):
    """hn_lane_loss (OHEM classification + masked Huber regression) against losses.lane_cls_loss /
    lane_reg_loss, which are pinned against the live reference (tests/test_cpu_losses.py).  Cases: the usual few positives, none
    at all (positive_num clamps to 1, one hard negative), more positives than negative_ratio leaves negatives for."""
    g = torch.Generator().manual_seed(B * 1000 + na + n_pos)
    gt_cls = torch.zeros((B, na, 2)); gt_cls[:, :, 0] = 1.0
    gt_loc = torch.zeros((B, na, L))
    for b in range(B):
        pos = torch.randperm(na, generator=g)[:n_pos]
        gt_cls[b, pos, 0], gt_cls[b, pos, 1] = 0.0, 1.0
        gt_loc[b, pos] = torch.randn((n_pos, L), generator=g) * 2.0
        gt_loc[b, pos, 7] = 0.0  # an invalid point inside a positive row
        gt_loc[b, pos, L - 2] = torch.randint(1, 80, (n_pos,), generator=g).float()
        gt_loc[b, pos, L - 1] = torch.randint(1, 80, (n_pos,), generator=g).float()
    gt_cls, gt_loc = gt_cls.cuda(), gt_loc.cuda()
    pc0 = (torch.randn((B, na, 2), generator=g) * 1.5).cuda()
    pl0 = (torch.randn((B, na, L), generator=g) * 1.5).cuda()
    res = {}
    for name in ("torch", "native"):
        pc, pl = pc0.clone().requires_grad_(), pl0.clone().requires_grad_()
        if name == "torch":
            pos, neg, pmask, npos = losses.lane_cls_loss(gt_cls, pc)
            loc = losses.lane_reg_loss(pmask, npos, gt_loc, pl)
        else:
            pos, neg, loc = losses.NativeLaneLoss.apply(gt_cls, pc, gt_loc, pl)
        (1.0 * pos + 2.0 * neg + 3.0 * loc).backward()
        res[name] = (float(pos), float(neg), float(loc), pc.grad.clone(), pl.grad.clone())
    a, b = res["native"], res["torch"]
    for i in range(3):
        assert abs(a[i] - b[i]) <= 3e-6 * max(abs(b[i]), 1e-3), (i, a[:3], b[:3])
    assert float((a[3] - b[3]).abs().max()) <= 2e-5 * max(float(b[3].abs().max()), 1e-6)
    assert float((a[4] - b[4]).abs().max()) <= 2e-5 * max(float(b[4].abs().max()), 1e-6)
    assert int(((a[3] != 0) != (b[3] != 0)).sum()) == 0, "a different set of hard negatives"
