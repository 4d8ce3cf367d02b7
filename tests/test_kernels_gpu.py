"""Per-kernel parity (-m gpu): every libhgk entry point, called through the C-ABI, against a
float64 CPU reference built from the same torch ops the oracle uses.  Tolerances are
fp32-class (1e-5 relative to the tensor's max) unless stated."""
import pytest
import torch
import torch.nn.functional as F

from hgk_testlib import (DEV, call, ptr, rnd, dev32, nhwc, from_nhwc, relerr, affine_act, pack_w, lib)

pytestmark = pytest.mark.gpu

CONV_SHAPES = [
    # N, H, W, Cin, Cout, k
    (2, 16, 16, 64, 128, 1),
    (2, 16, 16, 128, 64, 3),
    (3, 5, 7, 12, 20, 3),        # ragged: nothing a multiple of the tile sizes
    (3, 5, 7, 20, 12, 1),
    (2, 1, 1, 256, 128, 1),      # 1x1 neck of config 1
    (2, 1, 1, 64, 64, 3),
    (1, 64, 64, 256, 16, 1),     # out_conv
    (1, 64, 64, 16, 256, 1),     # in_conv
    (2, 8, 8, 256, 256, 3),
    (5, 4, 4, 128, 256, 1),
]


def _conv_ref(x, w, b, k):
    return F.conv2d(x, w, b, padding=k // 2)


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "full"])
def test_conv_fwd(shape, variant):
    N, H, W, Ci, Co, k = shape
    x = rnd("x", (N, Ci, H, W))
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    b = rnd("b", (Co,))
    full = variant == "full"
    xs, xt = (rnd("xs", (Ci,), 0.5, 1.5), rnd("xt", (Ci,), -0.3, 0.3)) if full else (None, None)
    res = rnd("res", (N, Co, H, W)) if full else None
    rs, rt = (rnd("rs", (Co,), 0.5, 1.5), rnd("rt", (Co,), -0.3, 0.3)) if full else (None, None)
    y0 = rnd("y0", (N, Co, H, W)) if full else None
    ref = _conv_ref(affine_act(x, xs, xt, True), w, b, k)
    if full:
        ref = ref + affine_act(res, rs, rt, True) + y0
    dx, dw, db = nhwc(x), pack_w(w, 0), dev32(b)
    dxs, dxt = (dev32(xs), dev32(xt)) if full else (None, None)
    dres = nhwc(res) if full else None
    drs, drt = (dev32(rs), dev32(rt)) if full else (None, None)
    y = nhwc(y0) if full else torch.empty(N, H, W, Co, device=DEV)
    ssum = torch.zeros(Co, device=DEV, dtype=torch.float64)
    ssq = torch.zeros(Co, device=DEV, dtype=torch.float64)
    call("conv_nhwc", ptr(dx), ptr(dxs), ptr(dxt), 1, N, H, W, Ci, ptr(dw), k, 0, ptr(db), Co,
         ptr(dres), ptr(drs), ptr(drt), 1, ptr(y), int(full), ptr(ssum), ptr(ssq), 1)
    torch.cuda.synchronize()
    assert relerr(from_nhwc(y), ref) < 1e-5
    assert relerr(ssum.cpu(), ref.sum(dim=(0, 2, 3))) < 1e-5
    assert relerr(ssq.cpu(), (ref * ref).sum(dim=(0, 2, 3))) < 1e-5


@pytest.mark.parametrize("shape", [(1, 64, 64, 256, 16, 1), (1, 64, 64, 16, 256, 1), (2, 5, 7, 128, 16, 1),
                                   (3, 5, 5, 16, 128, 1), (2, 9, 3, 256, 16, 1)])
@pytest.mark.parametrize("variant", ["plain", "full"])
def test_conv_skinny_heads(shape, variant):
    """out_conv / in_conv (16-channel side) without a statistics epilogue: the streaming kernels of conv_skinny.cu."""
    N, H, W, Ci, Co, k = shape
    x = rnd("x", (N, Ci, H, W))
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    b = rnd("b", (Co,))
    full = variant == "full"
    xs, xt = (rnd("xs", (Ci,), 0.5, 1.5), rnd("xt", (Ci,), -0.3, 0.3)) if full else (None, None)
    res = rnd("res", (N, Co, H, W)) if full else None
    rs, rt = (rnd("rs", (Co,), 0.5, 1.5), rnd("rt", (Co,), -0.3, 0.3)) if full else (None, None)
    y0 = rnd("y0", (N, Co, H, W)) if full else None
    ref = _conv_ref(affine_act(x, xs, xt, True), w, b, k)
    if full:
        ref = ref + affine_act(res, rs, rt, True) + y0
    dx, dw, db = nhwc(x), pack_w(w, 0), dev32(b)
    dxs, dxt = (dev32(xs), dev32(xt)) if full else (None, None)
    dres = nhwc(res) if full else None
    drs, drt = (dev32(rs), dev32(rt)) if full else (None, None)
    y = nhwc(y0) if full else torch.empty(N, H, W, Co, device=DEV)
    call("conv_nhwc", ptr(dx), ptr(dxs), ptr(dxt), 1, N, H, W, Ci, ptr(dw), k, 0, ptr(db) if full else 0, Co,
         ptr(dres), ptr(drs), ptr(drt), 1, ptr(y), int(full), 0, 0, 1)
    torch.cuda.synchronize()
    if not full:
        ref = ref - b.view(1, -1, 1, 1)
    assert relerr(from_nhwc(y), ref) < 1e-5


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_dgrad_and_wgrad(shape):
    N, H, W, Ci, Co, k = shape
    xs, xt = rnd("xs", (Ci,), 0.5, 1.5), rnd("xt", (Ci,), -0.3, 0.3)
    xin = rnd("x", (N, Ci, H, W))
    a = affine_act(xin, xs, xt, True).requires_grad_(True)
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2).requires_grad_(True)
    b = rnd("b", (Co,)).requires_grad_(True)
    dz = rnd("dz", (N, Co, H, W))
    _conv_ref(a, w, b, k).backward(dz)
    # data gradient: same kernel, flipped taps, [tap][Cout][Cin] weights, plus residual + accumulate
    extra = rnd("extra", (N, Ci, H, W))
    g0 = rnd("g0", (N, Ci, H, W))
    gx = nhwc(g0)
    ddz, dextra = nhwc(dz), nhwc(extra)
    wp = pack_w(w.detach(), 1)
    call("conv_nhwc", ptr(ddz), 0, 0, 0, N, H, W, Co, ptr(wp), k, 1, 0, Ci, ptr(dextra), 0, 0, 0, ptr(gx), 1, 0, 0, 1)
    torch.cuda.synchronize()
    assert relerr(from_nhwc(gx), a.grad + extra + g0) < 1e-5
    # weight / bias gradient written with OIHW strides, accumulating onto existing values
    w0, b0 = rnd("w0", (Co, Ci, k, k)), rnd("b0", (Co,))
    gw, gb = dev32(w0), dev32(b0)
    dxin, dxs, dxt = nhwc(xin), dev32(xs), dev32(xt)
    call("conv_wgrad_nhwc", ptr(dxin), ptr(dxs), ptr(dxt), 1, N, H, W, Ci, ptr(ddz), Co, k, ptr(gw), Ci * k * k, k * k, 1,
         ptr(gb))
    torch.cuda.synchronize()
    assert relerr(gw.cpu().double() - w0, w.grad) < 2e-5
    assert relerr(gb.cpu().double() - b0, b.grad) < 2e-5


def test_pack_weights():
    ws = [rnd("w%d" % i, s) for i, s in enumerate([(8, 4, 3, 3), (16, 8, 1, 1), (4, 12, 3, 3)])]
    flat = torch.cat([w.reshape(-1) for w in ws])
    src = dev32(flat)
    rows, off, so = [], 0, 0
    for w in ws:
        O, I, kh, kw = w.shape
        for mode in (0, 1):
            rows.append([so, off, O, I, kh * kw, mode])
            off += w.numel()
        so += w.numel()
    dst = torch.zeros(off, device=DEV)
    table = torch.tensor(rows, dtype=torch.long, device=DEV)
    call("pack_weights", ptr(src), ptr(dst), ptr(table), len(rows))
    torch.cuda.synchronize()
    o = 0
    for w in ws:
        for mode in (0, 1):
            got = dst[o:o + w.numel()].cpu()
            assert torch.equal(got, pack_w(w, mode).cpu().reshape(-1))
            o += w.numel()


@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 32, 48), (1, 256, 256)])
def test_stem_fwd_and_wgrad(shape):
    N, H, W = shape
    x = rnd("img", (N, 3, H, W), 0.0, 1.0)
    w = rnd("w", (64, 3, 7, 7), -0.1, 0.1).requires_grad_(True)
    b = rnd("b", (64,)).requires_grad_(True)
    ref = F.conv2d(x, w, b, stride=2, padding=3)
    y = torch.empty(N, H // 2, W // 2, 64, device=DEV)
    ssum = torch.zeros(64, device=DEV, dtype=torch.float64)
    ssq = torch.zeros(64, device=DEV, dtype=torch.float64)
    dimg, dw_, db_ = dev32(x), dev32(w.detach()), dev32(b.detach())
    call("stem_conv7_fwd", ptr(dimg), N, H, W, ptr(dw_), ptr(db_), 64, ptr(y), ptr(ssum), ptr(ssq))
    torch.cuda.synchronize()
    assert relerr(from_nhwc(y), ref) < 1e-5
    assert relerr(ssum.cpu(), ref.sum(dim=(0, 2, 3))) < 1e-5
    assert relerr(ssq.cpu(), (ref * ref).sum(dim=(0, 2, 3))) < 1e-5
    dz = rnd("dz", (N, 64, H // 2, W // 2))
    ref.backward(dz)
    gw = torch.zeros(64, 3, 7, 7, device=DEV)
    gb = torch.zeros(64, device=DEV)
    ddz = nhwc(dz)
    call("stem_conv7_wgrad", ptr(dimg), N, H, W, ptr(ddz), 64, ptr(gw), ptr(gb))
    torch.cuda.synchronize()
    assert relerr(gw, w.grad) < 2e-5
    assert relerr(gb, b.grad) < 2e-5
    # the same with bn1's BatchNorm-backward apply evaluated on load (hgk_stem_conv7_wgrad_bnapply) == hgk_bn_bwd_apply
    # followed by hgk_stem_conv7_wgrad on the dz it wrote
    OHW = (N, H // 2, W // 2, 64)
    g, zz = nhwc(rnd("g", (N, 64, H // 2, W // 2))), nhwc(rnd("zz", (N, 64, H // 2, W // 2), -2.0, 2.0))
    v = {n: dev32(rnd(n, (64,), lo, hi)) for n, lo, hi in (("sc", 0.5, 1.5), ("sh", -0.3, 0.3), ("mu", -0.2, 0.2),
                                                           ("cA", 0.5, 1.5), ("cB", -0.2, 0.2), ("cC", -0.1, 0.1))}
    dz_ref = g.clone()
    call("bn_bwd_apply", ptr(dz_ref), ptr(zz), ptr(v["sc"]), ptr(v["sh"]), 1, ptr(v["mu"]), ptr(v["cA"]), ptr(v["cB"]), ptr(v["cC"]),
         OHW[0] * OHW[1] * OHW[2], 64)
    gw_ref, gb_ref = torch.zeros(64, 3, 7, 7, device=DEV), torch.zeros(64, device=DEV)
    call("stem_conv7_wgrad", ptr(dimg), N, H, W, ptr(dz_ref), 64, ptr(gw_ref), ptr(gb_ref))
    gw2, gb2 = torch.zeros(64, 3, 7, 7, device=DEV), torch.zeros(64, device=DEV)
    call("stem_conv7_wgrad_bnapply", ptr(dimg), N, H, W, ptr(g), ptr(zz), ptr(v["sc"]), ptr(v["sh"]), 1, ptr(v["mu"]), ptr(v["cA"]),
         ptr(v["cB"]), ptr(v["cC"]), 64, ptr(gw2), ptr(gb2))
    torch.cuda.synchronize()
    assert relerr(gw2, gw_ref) < 2e-5
    assert relerr(gb2, gb_ref) < 2e-5


@pytest.mark.parametrize("shape", [(2, 8, 8, 64), (3, 5, 7, 12), (2, 1, 1, 128), (4, 16, 16, 256), (2, 4, 4, 320)])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_fwd_bwd(shape, training):
    N, H, W, C = shape
    z = rnd("z", (N, C, H, W), -2.0, 3.0).requires_grad_(True)
    gamma = rnd("gamma", (C,), 0.3, 1.2).requires_grad_(True)
    beta = rnd("beta", (C,), -0.3, 0.3).requires_grad_(True)
    rm, rv = rnd("rm", (C,), -0.2, 0.2), rnd("rv", (C,), 0.5, 1.5)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = F.relu(F.batch_norm(z, rm_ref, rv_ref, gamma, beta, training=training, momentum=0.1, eps=1e-5))
    dy = rnd("dy", (N, C, H, W))
    y.backward(dy)
    P = N * H * W
    dz_ = nhwc(z.detach())
    dgamma, dbeta, drm, drv = dev32(gamma.detach()), dev32(beta.detach()), dev32(rm), dev32(rv)
    scale, shift, mean, invstd = [torch.empty(C, device=DEV) for _ in range(4)]
    if training:
        ssum = z.detach().sum(dim=(0, 2, 3)).to(DEV)
        ssq = (z.detach() ** 2).sum(dim=(0, 2, 3)).to(DEV)
        call("bn_finalize", ptr(ssum), ptr(ssq), P, ptr(dgamma), ptr(dbeta), 1e-5, 0.1, ptr(drm), ptr(drv), ptr(scale),
             ptr(shift), ptr(mean), ptr(invstd), C)
        torch.cuda.synchronize()
        if P > 1:
            assert relerr(drm, rm_ref) < 1e-5 and relerr(drv, rv_ref) < 1e-5
    else:
        call("bn_eval_prepare", ptr(dgamma), ptr(dbeta), ptr(drm), ptr(drv), 1e-5, ptr(scale), ptr(shift), ptr(mean),
             ptr(invstd), C)
    # the consumers' on-load activation reproduces relu(bn(z))
    act = torch.empty(N, C, H, W, device=DEV)
    call("nhwc_to_nchw", ptr(dz_), ptr(scale), ptr(shift), 1, N, H, W, C, ptr(act))
    torch.cuda.synchronize()
    assert relerr(act, y) < (2e-4 if P <= 2 else 2e-5)     # 2 samples: xhat = +-1 exactly, ill-conditioned
    # backward: reduce -> finalize -> apply (in place)
    g = nhwc(dy)
    sg = torch.zeros(C, device=DEV, dtype=torch.float64)
    sgx = torch.zeros(C, device=DEV, dtype=torch.float64)
    gg0, gb0 = rnd("gg0", (C,)), rnd("gb0", (C,))
    ggamma, gbeta = dev32(gg0), dev32(gb0)
    cA, cB, cC = [torch.empty(C, device=DEV) for _ in range(3)]
    call("bn_bwd_reduce", ptr(g), ptr(dz_), ptr(scale), ptr(shift), 1, ptr(mean), ptr(invstd), P, C, ptr(sg), ptr(sgx))
    call("bn_bwd_finalize", ptr(sg), ptr(sgx), P, ptr(dgamma), ptr(mean), ptr(invstd), int(training), ptr(ggamma),
         ptr(gbeta), ptr(cA), ptr(cB), ptr(cC), C)
    call("bn_bwd_apply", ptr(g), ptr(dz_), ptr(scale), ptr(shift), 1, ptr(mean), ptr(cA), ptr(cB), ptr(cC), P, C)
    torch.cuda.synchronize()
    tol = 2e-4 if P <= 2 else 2e-5       # 2 samples: xhat = +-1, invstd ~ 1/|dz|: badly conditioned in any precision
    assert relerr(from_nhwc(g), z.grad) < tol
    assert relerr(ggamma.cpu().double() - gg0, gamma.grad) < 2e-5
    assert relerr(gbeta.cpu().double() - gb0, beta.grad) < 2e-5


@pytest.mark.parametrize("shape", [(2, 8, 8, 64), (3, 6, 10, 12), (1, 2, 2, 256)])
def test_maxpool_fwd_bwd(shape):
    N, H, W, C = shape
    s, t = rnd("s", (C,), 0.5, 1.5), rnd("t", (C,), -0.3, 0.3)
    x = rnd("x", (N, C, H, W))
    a = affine_act(x, s, t, True).requires_grad_(True)
    y = F.max_pool2d(a, 2, 2)
    dy = rnd("dy", y.shape)
    y.backward(dy)
    dx, ds, dt = nhwc(x), dev32(s), dev32(t)
    out = torch.empty(N, H // 2, W // 2, C, device=DEV)
    call("maxpool2_fwd", ptr(dx), ptr(ds), ptr(dt), 1, N, H, W, C, ptr(out))
    torch.cuda.synchronize()
    assert relerr(from_nhwc(out), y) < 1e-6
    g0 = rnd("g0", (N, C, H, W))
    gx = nhwc(g0)
    ddy = nhwc(dy)
    call("maxpool2_bwd", ptr(dx), ptr(ds), ptr(dt), 1, N, H, W, C, ptr(ddy), ptr(gx), 1)
    torch.cuda.synchronize()
    # ties happen only where relu clamps to 0; there the reference's ReLU' = 0 kills the gradient, so
    # compare the gradient w.r.t. the PRE-activation tensor
    mask = (a > 0).double()
    assert relerr((from_nhwc(gx) - g0) * mask, a.grad * mask) < 1e-6
    gx2 = torch.empty(N, H, W, C, device=DEV)
    call("maxpool2_bwd", ptr(dx), ptr(ds), ptr(dt), 1, N, H, W, C, ptr(ddy), ptr(gx2), 0)
    torch.cuda.synchronize()
    assert relerr(from_nhwc(gx2) * mask, a.grad * mask) < 1e-6


@pytest.mark.parametrize("up", [1, 0])
def test_add_upsample_fwd_bwd(up):
    N, H, W, C = 2, 8, 12, 64
    ha, wa = (H // 2, W // 2) if up else (H, W)
    a, b = rnd("a", (N, C, ha, wa)), rnd("b", (N, C, H, W))
    sa, ta, sb, tb = rnd("sa", (C,), 0.5, 1.5), rnd("ta", (C,), -0.3, 0.3), rnd("sb", (C,), 0.5, 1.5), rnd("tb", (C,), -0.3, 0.3)
    ra = affine_act(a, sa, ta, True)
    if up:
        ra = ra.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    ref = ra + affine_act(b, sb, tb, True)
    da, db_, dsa, dta, dsb, dtb = nhwc(a), nhwc(b), dev32(sa), dev32(ta), dev32(sb), dev32(tb)
    y = torch.empty(N, H, W, C, device=DEV)
    call("add_fwd", ptr(da), ptr(dsa), ptr(dta), 1, up, ptr(db_), ptr(dsb), ptr(dtb), 1, N, H, W, C, ptr(y))
    torch.cuda.synchronize()
    assert relerr(from_nhwc(y), ref) < 1e-6
    # plain operands
    call("add_fwd", ptr(da), 0, 0, 0, up, ptr(db_), 0, 0, 0, N, H, W, C, ptr(y))
    torch.cuda.synchronize()
    ra = a.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) if up else a
    assert relerr(from_nhwc(y), ra + b) < 1e-6
    if up:
        dy = rnd("dy", (N, C, H, W))
        g0 = rnd("g0", (N, C, ha, wa))
        ga = nhwc(g0)
        ddy = nhwc(dy)
        call("upsample2_bwd", ptr(ddy), N, H, W, C, ptr(ga), 1)
        torch.cuda.synchronize()
        ref_g = F.avg_pool2d(dy, 2) * 4 + g0
        assert relerr(from_nhwc(ga), ref_g) < 1e-6


@pytest.mark.parametrize("shape", [(2, 8, 8, 64), (3, 16, 32, 256), (24, 8, 8, 128)])
@pytest.mark.parametrize("which", ["maxpool", "upsample"])
def test_pool_upsample_bwd_with_fused_bn_reduction(shape, which):
    """hgk_maxpool2_bwd_bnred / hgk_upsample2_bwd_bnred == the plain kernel followed by hgk_bn_bwd_reduce_fin over the gradient it
    completed: same gradient bit for bit, same sums / dgamma / dbeta / cA / cB / cC to rounding, ticket re-armed."""
    N, H, W, C = shape
    v = {n: dev32(rnd(n, (C,), lo, hi)) for n, lo, hi in (("sc", 0.5, 1.5), ("sh", -0.3, 0.3), ("mu", -0.2, 0.2),
                                                           ("iv", 0.5, 2.0), ("gamma", 0.5, 1.5))}

    def fresh():
        return dict(sg=torch.zeros(C, device=DEV, dtype=torch.float64), sgx=torch.zeros(C, device=DEV, dtype=torch.float64),
                    dgamma=torch.zeros(C, device=DEV), dbeta=torch.zeros(C, device=DEV),
                    cA=torch.zeros(C, device=DEV), cB=torch.zeros(C, device=DEV), cC=torch.zeros(C, device=DEV))
    r, f = fresh(), fresh()
    ticket = torch.zeros(1, device=DEV, dtype=torch.int32)
    if which == "maxpool":
        x = nhwc(rnd("x", (N, C, H, W), -2.0, 2.0))                      # pooled tensor (pre-BN), its gradient has the same shape
        dy = nhwc(rnd("dy", (N, C, H // 2, W // 2)))
        g0 = nhwc(rnd("g0", (N, C, H, W)))
        g_ref, g_fus = g0.clone(), g0.clone()
        call("maxpool2_bwd", ptr(x), ptr(v["sc"]), ptr(v["sh"]), 1, N, H, W, C, ptr(dy), ptr(g_ref), 1)
        z, P = x, N * H * W
        call("maxpool2_bwd_bnred", ptr(x), ptr(v["sc"]), ptr(v["sh"]), 1, N, H, W, C, ptr(dy), ptr(g_fus), 1, ptr(v["mu"]), ptr(v["iv"]),
             ptr(f["sg"]), ptr(f["sgx"]), ptr(v["gamma"]), 1, ptr(f["dgamma"]), ptr(f["dbeta"]), ptr(f["cA"]), ptr(f["cB"]), ptr(f["cC"]),
             ptr(ticket))
    else:
        z = nhwc(rnd("z", (N, C, H // 2, W // 2), -2.0, 2.0))            # tensor the up-sampled branch came from
        dy = nhwc(rnd("dy", (N, C, H, W)))
        g0 = nhwc(rnd("g0", (N, C, H // 2, W // 2)))
        g_ref, g_fus = g0.clone(), g0.clone()
        call("upsample2_bwd", ptr(dy), N, H, W, C, ptr(g_ref), 1)
        P = N * (H // 2) * (W // 2)
        call("upsample2_bwd_bnred", ptr(dy), N, H, W, C, ptr(g_fus), 1, ptr(z), ptr(v["sc"]), ptr(v["sh"]), 1, ptr(v["mu"]), ptr(v["iv"]),
             ptr(f["sg"]), ptr(f["sgx"]), ptr(v["gamma"]), 1, ptr(f["dgamma"]), ptr(f["dbeta"]), ptr(f["cA"]), ptr(f["cB"]), ptr(f["cC"]),
             ptr(ticket))
    call("bn_bwd_reduce_fin", ptr(g_ref), ptr(z), ptr(v["sc"]), ptr(v["sh"]), 1, ptr(v["mu"]), ptr(v["iv"]), P, C, ptr(r["sg"]),
         ptr(r["sgx"]), ptr(v["gamma"]), 1, ptr(r["dgamma"]), ptr(r["dbeta"]), ptr(r["cA"]), ptr(r["cB"]), ptr(r["cC"]), ptr(ticket))
    torch.cuda.synchronize()
    assert torch.equal(g_fus, g_ref)
    assert int(ticket.item()) == 0
    for key in ("sg", "sgx", "dgamma", "dbeta", "cA", "cB", "cC"):
        assert relerr(f[key].float(), r[key].float()) < 1e-5, key


def test_add_into_and_layout():
    for n in (1024, 1030, 7):
        a, b = rnd("a", (n,)), rnd("b", (n,))
        da, db_ = dev32(a), dev32(b)
        call("add_into", ptr(da), ptr(db_), n, 1)
        torch.cuda.synchronize()
        assert relerr(db_, a + b) < 1e-6
        call("add_into", ptr(da), ptr(db_), n, 0)
        torch.cuda.synchronize()
        assert torch.equal(db_.cpu(), a.float())
    for (N, C, H, W) in [(2, 16, 64, 64), (3, 20, 5, 7), (1, 3, 33, 65), (2, 256, 4, 4)]:
        x = rnd("x", (N, C, H, W))
        dx = dev32(x)
        y = torch.empty(N, H, W, C, device=DEV)
        call("nchw_to_nhwc", ptr(dx), N, C, H, W, ptr(y))
        torch.cuda.synchronize()
        assert torch.equal(y.cpu(), x.float().permute(0, 2, 3, 1).contiguous())
        z = torch.empty(N, C, H, W, device=DEV)
        call("nhwc_to_nchw", ptr(y), 0, 0, 0, N, H, W, C, ptr(z))
        torch.cuda.synchronize()
        assert torch.equal(z.cpu(), x.float())


def test_avgpool_linear():
    N, H, W, C, k = 3, 4, 4, 64, 4
    s, t = rnd("s", (C,), 0.5, 1.5), rnd("t", (C,), -0.3, 0.3)
    x = rnd("x", (N, C, H, W))
    a = affine_act(x, s, t, True).requires_grad_(True)
    pooled = F.avg_pool2d(a, k)
    w = rnd("w", (7, C), -0.2, 0.2).requires_grad_(True)
    b = rnd("b", (7,)).requires_grad_(True)
    feat = pooled.view(N, C)
    out = F.linear(feat, w, b)
    dy = rnd("dy", (N, 7))
    feat.retain_grad()
    out.backward(dy)
    dx, ds, dt = nhwc(x), dev32(s), dev32(t)
    y = torch.empty(N, 1, 1, C, device=DEV)
    call("avgpool_fwd", ptr(dx), ptr(ds), ptr(dt), 1, N, H, W, C, k, ptr(y))
    o = torch.empty(N, 7, device=DEV)
    dw_, db_ = dev32(w.detach()), dev32(b.detach())
    call("linear_fwd", ptr(y), ptr(dw_), ptr(db_), N, C, 7, ptr(o))
    torch.cuda.synchronize()
    assert relerr(y.view(N, C), feat) < 1e-6
    assert relerr(o, out) < 1e-5
    gfeat = torch.empty(N, C, device=DEV)
    gw, gb = torch.zeros(7, C, device=DEV), torch.zeros(7, device=DEV)
    ddy = dev32(dy)
    call("linear_bwd", ptr(y), ptr(dw_), ptr(ddy), N, C, 7, ptr(gfeat), ptr(gw), ptr(gb))
    gx = torch.empty(N, H, W, C, device=DEV)
    call("avgpool_bwd", ptr(gfeat), N, H, W, C, k, ptr(gx), 0)
    torch.cuda.synchronize()
    assert relerr(gfeat, feat.grad) < 1e-5
    assert relerr(gw, w.grad) < 1e-5 and relerr(gb, b.grad) < 1e-5
    assert relerr(from_nhwc(gx), a.grad) < 1e-5


def test_mse_and_rmsprop_and_criterion():
    from oracle import hg_oracle as O
    o, t = rnd("o", (3, 16, 16, 16)), rnd("t", (3, 16, 16, 16), 0.0, 1.0)
    n = o.numel()
    do_, dt = dev32(o), dev32(t)
    g = torch.empty_like(do_)
    acc = torch.zeros(1, device=DEV, dtype=torch.float64)
    call("mse_fwd_bwd", ptr(do_), ptr(dt), n, 1.0 / n, 1.0, ptr(g), 0, ptr(acc))
    torch.cuda.synchronize()
    ref = O.mse_loss([o], t)
    assert abs(float(acc) - float(ref)) < 1e-6 * float(ref)
    assert relerr(g, 2 * (o - t) / n) < 1e-6
    # flat RMSprop == torch.optim.RMSprop semantics (oracle.rmsprop_step), two steps, odd length
    for n in (1000, 1003):
        p, gr = rnd("p", (n,)), rnd("g", (n,), -0.01, 0.01)
        v = torch.zeros(n, dtype=torch.float64)
        dp, dg, dv = dev32(p), dev32(gr), dev32(v)
        params, grads, sq = {"p": p.clone()}, {"p": gr}, {"p": v}
        for _ in range(2):
            call("rmsprop_flat", ptr(dp), ptr(dg), ptr(dv), n, 2.5e-4, 0.99, 1e-8, 1.0)
            O.rmsprop_step(params, grads, sq, lr=2.5e-4)
        torch.cuda.synchronize()
        assert relerr(dp, params["p"]) < 1e-6 and relerr(dv, sq["p"]) < 1e-5
    # grad_scale folds the 1/world_size of the data-parallel mean
    dp2, dv2 = dev32(p), dev32(v * 0)
    dg2 = dev32(gr * 4)
    call("rmsprop_flat", ptr(dp2), ptr(dg2), ptr(dv2), n, 2.5e-4, 0.99, 1e-8, 0.25)
    dp3, dv3 = dev32(p), dev32(v * 0)
    call("rmsprop_flat", ptr(dp3), ptr(dev32(gr)), ptr(dv3), n, 2.5e-4, 0.99, 1e-8, 1.0)
    torch.cuda.synchronize()
    assert relerr(dp2, dp3) < 1e-6
    # pylib/Criterion.py
    from pose_adv_aug_b200.pylib import Criterion as C
    pred = torch.sigmoid(rnd("cp", (2, 4, 8, 8)))
    gt = (rnd("cg", (2, 4, 8, 8)) > 0.5).double()
    wt = 1.0 + 4.0 * gt
    for fn_new, fn_ref in ((C.weighted_L2, O.weighted_L2), (C.weighted_sigmoid_crossentropy, O.weighted_sigmoid_crossentropy)):
        p1 = pred.clone().requires_grad_(True)
        l_ref = fn_ref(p1, gt, wt)
        l_ref.backward()
        p2 = dev32(pred).requires_grad_(True)
        l_new = fn_new(p2, dev32(gt), dev32(wt))
        (l_new * 3.0).backward()
        assert abs(float(l_new) - float(l_ref)) < 1e-5 * abs(float(l_ref))
        assert relerr(p2.grad, 3.0 * p1.grad) < 1e-5


def test_error_paths():
    from pose_adv_aug_b200._lib import get_lib
    L = get_lib()
    x = torch.zeros(4, device=DEV)
    s = torch.cuda.current_stream().cuda_stream
    rc = L.conv_nhwc(ptr(x), 0, 0, 0, 1, 1, 1, 6, ptr(x), 1, 0, 0, 4, 0, 0, 0, 0, ptr(x), 0, 0, 0, 1, s)
    assert rc == -1 and "multiples of 4" in L.last_error()
    rc = L.conv_nhwc(ptr(x), 0, 0, 0, 1, 1, 1, 4, ptr(x), 5, 0, 0, 4, 0, 0, 0, 0, ptr(x), 0, 0, 0, 1, s)
    assert rc == -1 and "ksize" in L.last_error()
    rc = L.maxpool2_fwd(ptr(x), 0, 0, 0, 1, 3, 3, 4, ptr(x), s)
    assert rc == -1 and "even" in L.last_error()
    assert L.cdll.hgk_device_ok() == 1


TC_SHAPES = [
    # N, H, W, Cin, Cout, k
    (2, 16, 16, 64, 128, 1),
    (2, 16, 16, 128, 64, 3),
    (3, 5, 7, 32, 64, 3),        # ragged pixel count (105 px: one partial tile)
    (2, 1, 1, 256, 128, 1),
    (1, 64, 64, 256, 256, 1),
    (2, 8, 8, 128, 128, 3),
    (5, 12, 12, 128, 256, 1),
    (1, 20, 24, 256, 128, 3),
    (8, 64, 64, 64, 128, 1),     # 256 tiles > 148 SMs: the two-CTAs-per-SM configuration
    (5, 64, 64, 128, 128, 3),    # 160 tiles, 3x3
    (5, 64, 64, 128, 256, 1),
    # 16x16-pixel image-tile kernel (conv_tc2.cu: H, W multiples of 16): every instantiation
    (1, 128, 128, 64, 64, 3),
    (1, 128, 128, 64, 64, 1),
    (2, 32, 32, 256, 128, 1),
    (3, 32, 48, 128, 128, 3),    # non-square, borders on every side of a tile
    (1, 16, 32, 64, 64, 1),
    (2, 16, 16, 256, 256, 1),
    (1, 48, 16, 128, 64, 1),
    # >= 148 tiles of 16x16: the two-halves-per-tile configuration (fewer tiles take the 16x8 configuration)
    (10, 64, 64, 128, 128, 3),
    (10, 64, 64, 128, 256, 1),
    (10, 64, 64, 64, 64, 3),
    (10, 64, 64, 256, 128, 1),
    # the same layer classes just off the 16-pixel grid: conv_tc_kernel, two-CTAs-per-SM configuration
    (6, 60, 60, 128, 128, 3),
    (8, 60, 60, 64, 128, 1),
]


def _pack_tc(w, mode, BN):
    """OIHW fp64 cpu -> (hi, lo) fp32 device arrays in the UMMA operand layout via hgk_pack_weights_tc."""
    O, I, kh, kw = w.shape
    taps = kh * kw
    N, K = (O, I) if mode == 0 else (I, O)
    src = dev32(w.reshape(-1))
    dst = torch.zeros(2 * w.numel(), device=DEV)
    table = torch.tensor([[0, 0, w.numel(), N, K, taps, mode, BN]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
    torch.cuda.synchronize()
    hi, lo = dst[:w.numel()], dst[w.numel():]
    # hi + lo reproduces w exactly in fp32 and hi is a TF32 number
    assert torch.equal((hi.double() + lo.double()).float().sort().values, src.sort().values)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    return hi, lo


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "full"])
def test_conv_tc_fwd_3xtf32(shape, variant):
    N, H, W, Ci, Co, k = shape
    x = rnd("x", (N, Ci, H, W))
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    b = rnd("b", (Co,))
    full = variant == "full"
    xs, xt = (rnd("xs", (Ci,), 0.5, 1.5), rnd("xt", (Ci,), -0.3, 0.3)) if full else (None, None)
    res = rnd("res", (N, Co, H, W)) if full else None
    rs, rt = (rnd("rs", (Co,), 0.5, 1.5), rnd("rt", (Co,), -0.3, 0.3)) if full else (None, None)
    y0 = rnd("y0", (N, Co, H, W)) if full else None
    ref = _conv_ref(affine_act(x, xs, xt, True), w, b, k)
    if full:
        ref = ref + affine_act(res, rs, rt, True) + y0
    hi, lo = _pack_tc(w, 0, Co)
    dx, db = nhwc(x), dev32(b)
    dxs, dxt = (dev32(xs), dev32(xt)) if full else (None, None)
    dres = nhwc(res) if full else None
    drs, drt = (dev32(rs), dev32(rt)) if full else (None, None)
    y = nhwc(y0) if full else torch.empty(N, H, W, Co, device=DEV)
    ssum = torch.zeros(Co, device=DEV, dtype=torch.float64)
    ssq = torch.zeros(Co, device=DEV, dtype=torch.float64)
    call("conv_tc_nhwc", ptr(dx), ptr(dxs), ptr(dxt), 1, N, H, W, Ci, ptr(hi), ptr(lo), k, ptr(db), Co,
         ptr(dres), ptr(drs), ptr(drt), 1, ptr(y), int(full), ptr(ssum), ptr(ssq))
    torch.cuda.synchronize()
    # 3xTF32: dropped lo*lo term ~2^-22 per product -> fp32-class result.  (The tensor core accumulates in
    # fp32 with truncation; the kernel splits the chain over several TMEM accumulators to keep that bias
    # at the 1e-6 level even at K = 9*256.)
    assert relerr(from_nhwc(y), ref) < 2e-5
    assert relerr(ssum.cpu(), ref.sum(dim=(0, 2, 3))) < 2e-5
    assert relerr(ssq.cpu(), (ref * ref).sum(dim=(0, 2, 3))) < 2e-5


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 128, 1), (3, 5, 7, 32, 64, 3), (2, 8, 8, 128, 128, 3),
                                   (5, 64, 64, 128, 256, 1), (3, 32, 48, 128, 128, 3), (6, 60, 60, 128, 128, 3),
                                   (10, 64, 64, 128, 128, 3)])
def test_conv_tc_fused_bn_finalize(shape):
    """conv + batch statistics + BatchNorm finaliser in ONE launch (last-CTA ticket): scale/shift/mean/invstd and the
    running statistics must equal nn.BatchNorm2d's on the convolution output; the ticket re-arms itself."""
    N, H, W, Ci, Co, k = shape
    x = rnd("x", (N, Ci, H, W))
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    b = rnd("b", (Co,))
    gamma, beta = rnd("gamma", (Co,), 0.5, 1.5), rnd("beta", (Co,), -0.3, 0.3)
    rm0, rv0 = rnd("rm", (Co,), -0.2, 0.2), rnd("rv", (Co,), 0.5, 1.5)
    ref = _conv_ref(x, w, b, k)
    cnt = N * H * W
    mean = ref.mean(dim=(0, 2, 3))
    var = ref.var(dim=(0, 2, 3), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    hi, lo = _pack_tc(w, 0, Co)
    dx, db = nhwc(x), dev32(b)
    dg, dbeta, drm, drv = dev32(gamma), dev32(beta), dev32(rm0), dev32(rv0)
    y = torch.empty(N, H, W, Co, device=DEV)
    sc, sh, sm, si = (torch.zeros(Co, device=DEV) for _ in range(4))
    ticket = torch.zeros(1, device=DEV, dtype=torch.int32)
    for rep in range(2):        # the second launch checks that the ticket was re-armed
        ssum = torch.zeros(Co, device=DEV, dtype=torch.float64)
        ssq = torch.zeros(Co, device=DEV, dtype=torch.float64)
        call("conv_tc_bn_nhwc", ptr(dx), 0, 0, 0, N, H, W, Ci, ptr(hi), ptr(lo), k, ptr(db), Co, 0, 0, 0, 0, ptr(y), 0,
             ptr(ssum), ptr(ssq), ptr(dg), ptr(dbeta), 1e-5, 0.1, ptr(drm), ptr(drv), ptr(sc), ptr(sh), ptr(sm), ptr(si),
             ptr(ticket))
        torch.cuda.synchronize()
        assert int(ticket.item()) == 0
        assert relerr(from_nhwc(y), ref) < 2e-5
        assert relerr(sm.cpu(), mean) < 2e-5 and relerr(si.cpu(), invstd) < 2e-5
        assert relerr(sc.cpu(), gamma * invstd) < 2e-5
        assert relerr(sh.cpu(), beta - mean * gamma * invstd) < 5e-5
    unb = var * cnt / (cnt - 1)
    rm1 = 0.9 * (0.9 * rm0 + 0.1 * mean) + 0.1 * mean          # two momentum updates
    rv1 = 0.9 * (0.9 * rv0 + 0.1 * unb) + 0.1 * unb
    assert relerr(drm.cpu(), rm1) < 2e-5 and relerr(drv.cpu(), rv1) < 2e-5


@pytest.mark.parametrize("shape", [(8, 16, 16, 128, 128, 3), (3, 32, 48, 128, 128, 3), (2, 64, 64, 64, 64, 3),
                                   (10, 64, 64, 128, 128, 3), (2, 128, 128, 64, 64, 3), (8, 16, 16, 256, 128, 3),
                                   (5, 64, 64, 128, 256, 1), (10, 64, 64, 256, 128, 1), (2, 32, 48, 256, 256, 1),
                                   (3, 20, 28, 64, 128, 1)])
@pytest.mark.parametrize("variant", ["plain", "full"])
def test_conv_tc_fwd_tf32_plus_2xbf16(shape, variant):
    """hgk_conv_tc_bn_x2_nhwc: x*w ~= xh*wh (TF32) + bf16(xl)*bf16(wh) + bf16(xh)*bf16(wl) on the image-tile kernel (3x3) and the
    persistent kernel (1x1, incl. a ragged last tile) with the weights in pack mode 2
    against the fp64 convolution: same tolerance as the 3xTF32 kernel (error ~3 * 2^-20 per product), BN+ReLU on load,
    shortcut, accumulate, fused BatchNorm statistics / finaliser."""
    N, H, W, Ci, Co, k = shape
    assert lib().cdll.hgk_conv_tc_x2_supported(N, H, W, Ci, Co, k) == 1
    x = rnd("x", (N, Ci, H, W))
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    b = rnd("b", (Co,))
    full = variant == "full"
    xs, xt = (rnd("xs", (Ci,), 0.5, 1.5), rnd("xt", (Ci,), -0.3, 0.3)) if full else (None, None)
    res = rnd("res", (N, Co, H, W)) if full else None
    rs, rt = (rnd("rs", (Co,), 0.5, 1.5), rnd("rt", (Co,), -0.3, 0.3)) if full else (None, None)
    y0 = rnd("y0", (N, Co, H, W)) if full else None
    ref = _conv_ref(affine_act(x, xs, xt, True), w, b, k)
    if full:
        ref = ref + affine_act(res, rs, rt, True) + y0
    gamma, beta = rnd("gamma", (Co,), 0.5, 1.5), rnd("beta", (Co,), -0.3, 0.3)
    rm0, rv0 = rnd("rm", (Co,), -0.2, 0.2), rnd("rv", (Co,), 0.5, 1.5)
    # pack mode 2: hi as the forward operand, "lo" buffer = bf16 cross-term operands
    src = dev32(w.reshape(-1))
    dst = torch.zeros(2 * w.numel(), device=DEV)
    table = torch.tensor([[0, 0, w.numel(), Co, Ci, k * k, 2, Co]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
    hi, x2 = dst[:w.numel()], dst[w.numel():]
    hi0, _ = _pack_tc(w, 0, Co)
    assert torch.equal(hi, hi0)
    dx, db = nhwc(x), dev32(b)
    dxs, dxt = (dev32(xs), dev32(xt)) if full else (None, None)
    dres = nhwc(res) if full else None
    drs, drt = (dev32(rs), dev32(rt)) if full else (None, None)
    y = nhwc(y0) if full else torch.empty(N, H, W, Co, device=DEV)
    dg, dbeta, drm, drv = dev32(gamma), dev32(beta), dev32(rm0), dev32(rv0)
    sc, sh, sm, si = (torch.zeros(Co, device=DEV) for _ in range(4))
    ticket = torch.zeros(1, device=DEV, dtype=torch.int32)
    ssum = torch.zeros(Co, device=DEV, dtype=torch.float64)
    ssq = torch.zeros(Co, device=DEV, dtype=torch.float64)
    call("conv_tc_bn_x2_nhwc", ptr(dx), ptr(dxs), ptr(dxt), 1, N, H, W, Ci, ptr(hi), ptr(x2), k, ptr(db), Co,
         ptr(dres), ptr(drs), ptr(drt), 1, ptr(y), int(full), ptr(ssum), ptr(ssq), ptr(dg), ptr(dbeta), 1e-5, 0.1, ptr(drm),
         ptr(drv), ptr(sc), ptr(sh), ptr(sm), ptr(si), ptr(ticket))
    torch.cuda.synchronize()
    assert int(ticket.item()) == 0
    err = relerr(from_nhwc(y), ref)
    print("TF32 + 2xBF16 %s %s: rel-to-max error %.2e" % (shape, variant, err))
    assert err < 2e-5
    assert relerr(ssum.cpu(), ref.sum(dim=(0, 2, 3))) < 2e-5
    assert relerr(ssq.cpu(), (ref * ref).sum(dim=(0, 2, 3))) < 2e-5
    mean = ref.mean(dim=(0, 2, 3))
    invstd = 1.0 / torch.sqrt(ref.var(dim=(0, 2, 3), unbiased=False) + 1e-5)
    assert relerr(sm.cpu(), mean) < 2e-5 and relerr(si.cpu(), invstd) < 2e-5


@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 64, 3), (2, 8, 8, 128, 128, 3), (1, 64, 64, 128, 64, 1)])
def test_bn_bwd_fused_finalizers(shape):
    """dgrad + BN-backward sums + finaliser in one launch, and bn_bwd_reduce + finaliser in one launch, both against
    the separate bn_bwd_reduce -> bn_bwd_finalize kernels."""
    N, H, W, Ci, Co, k = shape
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    dz = rnd("dz", (N, Co, H, W))
    bz = rnd("bz", (N, Ci, H, W))
    gamma = rnd("gamma", (Ci,), 0.5, 1.5)
    bsc, bsh = rnd("bsc", (Ci,), 0.5, 1.5), rnd("bsh", (Ci,), -0.3, 0.3)
    bmu, biv = rnd("bmu", (Ci,), -0.2, 0.2), rnd("biv", (Ci,), 0.5, 2.0)
    ddz, dbz = nhwc(dz), nhwc(bz)
    dgam, dbsc, dbsh, dbmu, dbiv = dev32(gamma), dev32(bsc), dev32(bsh), dev32(bmu), dev32(biv)
    hi, _ = _pack_tc(w, 1, Ci)
    P = N * H * W

    def fresh():
        return dict(sg=torch.zeros(Ci, device=DEV, dtype=torch.float64), sgx=torch.zeros(Ci, device=DEV, dtype=torch.float64),
                    dgamma=torch.full((Ci,), 0.25, device=DEV), dbeta=torch.full((Ci,), -0.5, device=DEV),
                    cA=torch.zeros(Ci, device=DEV), cB=torch.zeros(Ci, device=DEV), cC=torch.zeros(Ci, device=DEV))
    # reference chain: dgrad+bnstats -> bn_bwd_finalize
    r = fresh()
    gx_ref = torch.empty(N, H, W, Ci, device=DEV)
    call("conv_tc_dgrad_bnstats_nhwc", ptr(ddz), N, H, W, Co, ptr(hi), 0, k, Ci, 0, ptr(gx_ref), 0, ptr(dbz), ptr(dbsc), ptr(dbsh), 1,
         ptr(dbmu), ptr(dbiv), ptr(r["sg"]), ptr(r["sgx"]))
    call("bn_bwd_finalize", ptr(r["sg"]), ptr(r["sgx"]), P, ptr(dgam), ptr(dbmu), ptr(dbiv), 1, ptr(r["dgamma"]), ptr(r["dbeta"]),
         ptr(r["cA"]), ptr(r["cB"]), ptr(r["cC"]), Ci)
    torch.cuda.synchronize()
    ticket = torch.zeros(1, device=DEV, dtype=torch.int32)
    # (a) fused in the data-gradient kernel
    f = fresh()
    gx = torch.empty(N, H, W, Ci, device=DEV)
    call("conv_tc_dgrad_bnfin_nhwc", ptr(ddz), N, H, W, Co, ptr(hi), 0, k, Ci, 0, ptr(gx), 0, ptr(dbz), ptr(dbsc), ptr(dbsh), 1,
         ptr(dbmu), ptr(dbiv), ptr(f["sg"]), ptr(f["sgx"]), ptr(dgam), 1, ptr(f["dgamma"]), ptr(f["dbeta"]), ptr(f["cA"]),
         ptr(f["cB"]), ptr(f["cC"]), ptr(ticket))
    torch.cuda.synchronize()
    assert int(ticket.item()) == 0 and torch.equal(gx, gx_ref)
    for key in ("dgamma", "dbeta", "cA", "cB", "cC"):
        assert relerr(f[key], r[key]) < 1e-5, key
    # (b) fused in the stand-alone reduction (gx_ref plays dY)
    f2 = fresh()
    call("bn_bwd_reduce_fin", ptr(gx_ref), ptr(dbz), ptr(dbsc), ptr(dbsh), 1, ptr(dbmu), ptr(dbiv), P, Ci, ptr(f2["sg"]),
         ptr(f2["sgx"]), ptr(dgam), 1, ptr(f2["dgamma"]), ptr(f2["dbeta"]), ptr(f2["cA"]), ptr(f2["cB"]), ptr(f2["cC"]), ptr(ticket))
    torch.cuda.synchronize()
    assert int(ticket.item()) == 0
    for key in ("dgamma", "dbeta", "cA", "cB", "cC"):
        assert relerr(f2[key], r[key]) < 1e-5, key


@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 64, 3), (1, 64, 64, 128, 128, 3), (2, 32, 48, 64, 128, 1),
                                   (1, 16, 32, 256, 256, 1), (1, 128, 128, 64, 64, 3), (10, 64, 64, 128, 128, 3),
                                   (10, 64, 64, 256, 128, 1), (8, 16, 16, 128, 128, 3), (4, 16, 32, 256, 256, 1)])
@pytest.mark.parametrize("with_red", [0, 1])
def test_conv_tc_dgrad_bnapply(shape, with_red):
    """BatchNorm-backward apply evaluated on load by the image-tile data-gradient kernel == bn_bwd_apply followed by the
    plain data-gradient kernel; the dz side output equals bn_bwd_apply's result exactly."""
    N, H, W, Ci, Co, k = shape           # conv Ci -> Co; g, gz live on the Co side
    # "supported" is the planner's answer: layers of at most 12 tiles of 128 pixels take the cluster split-K kernel (no apply on
    # load) for their plain data gradient.  The fused kernel still accepts them; the reference chain below then runs a
    # different kernel (other summation order), so the data gradient is compared to rounding instead of bit for bit.
    same_kernel = lib().conv_tc_bnapply_supported(N, H, W, Co, Ci, k)
    assert same_kernel or N * H * W <= 12 * 128
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    g, gz = nhwc(rnd("g", (N, Co, H, W))), nhwc(rnd("gz", (N, Co, H, W)))
    v = {n: dev32(rnd(n, (Co,), lo, hi)) for n, lo, hi in (("sc", 0.5, 1.5), ("sh", -0.3, 0.3), ("mu", -0.2, 0.2),
                                                           ("cA", 0.5, 1.5), ("cB", -0.2, 0.2), ("cC", -0.1, 0.1))}
    hi, _ = _pack_tc(w, 1, Ci)
    extra = nhwc(rnd("extra", (N, Ci, H, W)))
    P = N * H * W
    # reference chain
    dz_ref = g.clone()
    call("bn_bwd_apply", ptr(dz_ref), ptr(gz), ptr(v["sc"]), ptr(v["sh"]), 1, ptr(v["mu"]), ptr(v["cA"]), ptr(v["cB"]), ptr(v["cC"]), P, Co)
    gx_ref = torch.empty(N, H, W, Ci, device=DEV)
    bz = nhwc(rnd("bz", (N, Ci, H, W)))
    b = {n: dev32(rnd(n, (Ci,), lo, hi)) for n, lo, hi in (("bsc", 0.5, 1.5), ("bsh", -0.3, 0.3), ("bmu", -0.2, 0.2),
                                                           ("biv", 0.5, 2.0), ("gamma", 0.5, 1.5))}

    def fresh():
        return dict(sg=torch.zeros(Ci, device=DEV, dtype=torch.float64), sgx=torch.zeros(Ci, device=DEV, dtype=torch.float64),
                    dgamma=torch.zeros(Ci, device=DEV), dbeta=torch.zeros(Ci, device=DEV),
                    cA=torch.zeros(Ci, device=DEV), cB=torch.zeros(Ci, device=DEV), cC=torch.zeros(Ci, device=DEV))
    r, f = fresh(), fresh()
    ticket = torch.zeros(1, device=DEV, dtype=torch.int32)
    if with_red:
        call("conv_tc_dgrad_bnfin_nhwc", ptr(dz_ref), N, H, W, Co, ptr(hi), 0, k, Ci, ptr(extra), ptr(gx_ref), 0, ptr(bz), ptr(b["bsc"]),
             ptr(b["bsh"]), 1, ptr(b["bmu"]), ptr(b["biv"]), ptr(r["sg"]), ptr(r["sgx"]), ptr(b["gamma"]), 1, ptr(r["dgamma"]),
             ptr(r["dbeta"]), ptr(r["cA"]), ptr(r["cB"]), ptr(r["cC"]), ptr(ticket))
    else:
        call("conv_tc_nhwc", ptr(dz_ref), 0, 0, 0, N, H, W, Co, ptr(hi), 0, k, 0, Ci, ptr(extra), 0, 0, 0, ptr(gx_ref), 0, 0, 0)
    # fused
    dz = torch.full((N, H, W, Co), 7.0, device=DEV)
    gx = torch.empty(N, H, W, Ci, device=DEV)
    red = ([ptr(bz), ptr(b["bsc"]), ptr(b["bsh"]), 1, ptr(b["bmu"]), ptr(b["biv"]), ptr(f["sg"]), ptr(f["sgx"]), ptr(b["gamma"]), 1,
            ptr(f["dgamma"]), ptr(f["dbeta"]), ptr(f["cA"]), ptr(f["cB"]), ptr(f["cC"]), ptr(ticket)] if with_red else [0] * 16)
    call("conv_tc_dgrad_bnapply_nhwc", ptr(g), ptr(gz), ptr(v["sc"]), ptr(v["sh"]), 1, ptr(v["mu"]), ptr(v["cA"]), ptr(v["cB"]),
         ptr(v["cC"]), ptr(dz), N, H, W, Co, ptr(hi), k, Ci, ptr(extra), ptr(gx), 0, *red)
    torch.cuda.synchronize()
    assert torch.equal(dz, dz_ref)               # same fp32 expression, every pixel written exactly once
    if same_kernel:
        assert torch.equal(gx, gx_ref)           # same operand values -> bit-identical accumulation
    else:
        assert relerr(gx, gx_ref) < 1e-5
    if with_red:
        assert int(ticket.item()) == 0
        for key in ("dgamma", "dbeta", "cA", "cB", "cC"):
            assert relerr(f[key], r[key]) < 1e-5, key


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tc_dgrad_1xtf32(shape):
    N, H, W, Ci, Co, k = shape
    a = rnd("x", (N, Ci, H, W)).requires_grad_(True)
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    dz = rnd("dz", (N, Co, H, W))
    _conv_ref(a, w, None, k).backward(dz)
    if Ci not in (64, 128, 256) or Co % 32:
        pytest.skip("data-gradient shape not covered by the tcgen05 path")
    extra, g0 = rnd("extra", (N, Ci, H, W)), rnd("g0", (N, Ci, H, W))
    gx, ddz, dextra = nhwc(g0), nhwc(dz), nhwc(extra)
    hi, _ = _pack_tc(w, 1, Ci)
    call("conv_tc_nhwc", ptr(ddz), 0, 0, 0, N, H, W, Co, ptr(hi), 0, k, 0, Ci, ptr(dextra), 0, 0, 0, ptr(gx), 1, 0, 0)
    torch.cuda.synchronize()
    # plain TF32 operands (10-bit mantissa): ~1e-3 relative
    assert relerr(from_nhwc(gx) - extra - g0, a.grad) < 3e-3


@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 64, 3), (3, 32, 48, 128, 128, 3), (2, 32, 32, 128, 256, 1),
                                   (1, 64, 64, 64, 64, 3), (2, 8, 8, 128, 128, 3), (3, 5, 7, 64, 64, 3),
                                   (1, 16, 32, 256, 256, 1), (1, 64, 64, 128, 64, 1)])
@pytest.mark.parametrize("acc", [0, 1])
def test_conv_tc_dgrad_bnstats(shape, acc):
    """Data gradient + fused BN-backward reduction (sum g, sum g*xhat) of the producer's BatchNorm."""
    N, H, W, Ci, Co, k = shape           # the convolution maps Ci -> Co; the gradient flows Co -> Ci
    a = rnd("x", (N, Ci, H, W)).requires_grad_(True)
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2)
    dz = rnd("dz", (N, Co, H, W))
    _conv_ref(a, w, None, k).backward(dz)
    extra, g0 = rnd("extra", (N, Ci, H, W)), rnd("g0", (N, Ci, H, W))
    bz = rnd("bz", (N, Ci, H, W))
    bsc, bsh = rnd("bsc", (Ci,), 0.5, 1.5), rnd("bsh", (Ci,), -0.3, 0.3)
    bmu, biv = rnd("bmu", (Ci,), -0.2, 0.2), rnd("biv", (Ci,), 0.5, 2.0)
    dy_ref = a.grad + extra + (g0 if acc else 0)
    mask = (bz * bsc.view(1, -1, 1, 1) + bsh.view(1, -1, 1, 1) > 0).double()
    g = dy_ref * mask
    xhat = (bz - bmu.view(1, -1, 1, 1)) * biv.view(1, -1, 1, 1)
    gx, ddz, dextra, dbz = nhwc(g0), nhwc(dz), nhwc(extra), nhwc(bz)
    hi, _ = _pack_tc(w, 1, Ci)
    sg = torch.zeros(Ci, device=DEV, dtype=torch.float64)
    sgx = torch.zeros(Ci, device=DEV, dtype=torch.float64)
    dbsc, dbsh, dbmu, dbiv = dev32(bsc), dev32(bsh), dev32(bmu), dev32(biv)
    call("conv_tc_dgrad_bnstats_nhwc", ptr(ddz), N, H, W, Co, ptr(hi), 0, k, Ci, ptr(dextra), ptr(gx), acc,
         ptr(dbz), ptr(dbsc), ptr(dbsh), 1, ptr(dbmu), ptr(dbiv), ptr(sg), ptr(sgx))
    torch.cuda.synchronize()
    assert relerr(from_nhwc(gx) - extra - (g0 if acc else 0), a.grad) < 3e-3      # plain TF32 operands
    # the sums are taken over the kernel's own (TF32-accurate) dy: compare against sums of the exact gradient
    # relative to the sum of magnitudes
    scale_g = g.abs().sum(dim=(0, 2, 3))
    assert float(((sg.cpu() - g.sum(dim=(0, 2, 3))).abs() / scale_g).max()) < 2e-3
    scale_gx = (g * xhat).abs().sum(dim=(0, 2, 3))
    assert float(((sgx.cpu() - (g * xhat).sum(dim=(0, 2, 3))).abs() / scale_gx).max()) < 2e-3


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 128, 1), (2, 16, 16, 128, 64, 3), (3, 5, 7, 64, 20, 3),
                                   (2, 1, 1, 256, 128, 1), (1, 64, 64, 256, 256, 1), (2, 8, 8, 128, 128, 3),
                                   (5, 12, 12, 128, 256, 1), (1, 20, 24, 256, 128, 3), (1, 64, 64, 256, 16, 1),
                                   (4, 64, 64, 128, 128, 3), (2, 32, 32, 128, 256, 1), (1, 128, 128, 64, 64, 3),
                                   (1, 64, 64, 64, 128, 1), (2, 16, 16, 256, 128, 1), (3, 16, 8, 64, 64, 1),
                                   (3, 8, 32, 64, 128, 3), (1, 64, 64, 128, 64, 3), (7, 4, 4, 128, 256, 1)])
def test_conv_wgrad_tc(shape):
    N, H, W, Ci, Co, k = shape
    xs, xt = rnd("xs", (Ci,), 0.5, 1.5), rnd("xt", (Ci,), -0.3, 0.3)
    xin = rnd("x", (N, Ci, H, W))
    a = affine_act(xin, xs, xt, True)
    w = rnd("w", (Co, Ci, k, k), -0.2, 0.2).requires_grad_(True)
    b = rnd("b", (Co,)).requires_grad_(True)
    dz = rnd("dz", (N, Co, H, W))
    _conv_ref(a, w, b, k).backward(dz)
    w0, b0 = rnd("w0", (k * k, Co, Ci)), rnd("b0", (Co,))
    gw, gb = dev32(w0), dev32(b0)
    dxin, dxs, dxt, ddz = nhwc(xin), dev32(xs), dev32(xt), nhwc(dz)
    call("conv_wgrad_tc_nhwc", ptr(dxin), ptr(dxs), ptr(dxt), 1, N, H, W, Ci, ptr(ddz), Co, k, ptr(gw), ptr(gb))
    torch.cuda.synchronize()
    ref_tap_major = w.grad.reshape(Co, Ci, k * k).permute(2, 0, 1)
    assert relerr(gw.cpu().double() - w0, ref_tap_major) < 3e-3          # plain TF32 operands
    assert relerr(gb.cpu().double() - b0, b.grad) < 2e-5                 # bias sums stay fp32
    # tap-major scratch -> OIHW .grad accumulation
    dst0 = rnd("dst0", (Co, Ci, k, k))
    dst = dev32(dst0)
    src = (gw.cpu().double() - w0).float().to(DEV).contiguous()
    table = torch.tensor([[0, 0, Co, Ci, k * k]], dtype=torch.long, device=DEV)
    call("unpack_add_grads", ptr(src), ptr(dst), ptr(table), 1)
    torch.cuda.synchronize()
    assert relerr(dst.cpu().double() - dst0, w.grad) < 3e-3


@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 32, 96), (2, 256, 256)])
@pytest.mark.parametrize("x2", [False, True])
def test_stem_tensor_core_space_to_depth_fwd(shape, x2):
    """nn.Conv2d(3,64,7,2,3) (ref models/asn_stacked_hg.py:223) on the tensor cores: hgk_stem_s2d_image + hgk_stem_s2d_weight
    turn it into a 4x4-tap stride-1 convolution over the 16-channel half-resolution image, run by the image-tile kernel
    (ksize = 4) with 3xTF32 or TF32 + 2xBF16 products, bias, batch statistics and the fused BatchNorm finaliser."""
    N, H, W = shape
    x = rnd("img", (N, 3, H, W), 0.0, 1.0)
    w = rnd("w", (64, 3, 7, 7), -0.1, 0.1)
    b = rnd("b", (64,))
    ref = F.conv2d(x, w, b, stride=2, padding=3)
    H2, W2 = H // 2, W // 2
    dimg, dw_, db_ = dev32(x), dev32(w), dev32(b)
    xs = torch.full((N, H2, W2, 16), float("nan"), device=DEV)
    ws = torch.full((64, 32, 4, 4), float("nan"), device=DEV)
    call("stem_s2d_image", ptr(dimg), N, H, W, ptr(xs))
    call("stem_s2d_weight", ptr(dw_), 64, ptr(ws))
    torch.cuda.synchronize()
    # the rearrangements themselves: exact
    want = torch.zeros(N, H2, W2, 16)
    for dy in range(2):
        for dx in range(2):
            for c in range(3):
                want[..., (dy * 2 + dx) * 3 + c] = x[:, c, dy::2, dx::2].float()
    assert torch.equal(xs.cpu(), want)
    wwant = torch.zeros(64, 32, 4, 4)
    for ty in range(4):
        for tx in range(4):
            for dy in range(2):
                for dx in range(2):
                    ky, kx = 2 * ty + dy - 1, 2 * tx + dx - 1
                    if 0 <= ky < 7 and 0 <= kx < 7:
                        for c in range(3):
                            wwant[:, (dy * 2 + dx) * 3 + c, ty, tx] = w[:, c, ky, kx].float()
    assert torch.equal(ws.cpu(), wwant)
    # the 4x4-tap form equals the 7x7 stride-2 convolution (fp64 identity check of the mapping)
    ident = F.conv2d(F.pad(want.permute(0, 3, 1, 2).double(), (2, 1, 2, 1)), wwant[:, :16].double(), b)
    assert relerr(ident, ref) < 1e-6
    dst = torch.zeros(2 * ws.numel(), device=DEV)
    table = torch.tensor([[0, 0, ws.numel(), 64, 32, 16, 2 if x2 else 0, 64]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(ws), ptr(dst), ptr(table), 1)
    hi, lo = dst[:ws.numel()], dst[ws.numel():]
    gamma, beta = rnd("gamma", (64,), 0.5, 1.5), rnd("beta", (64,), -0.3, 0.3)
    dg, dbeta, drm, drv = dev32(gamma), dev32(beta), torch.zeros(64, device=DEV), torch.ones(64, device=DEV)
    sc, sh, sm, si = (torch.zeros(64, device=DEV) for _ in range(4))
    ticket = torch.zeros(1, device=DEV, dtype=torch.int32)
    ssum = torch.zeros(64, device=DEV, dtype=torch.float64)
    ssq = torch.zeros(64, device=DEV, dtype=torch.float64)
    y = torch.full((N, H2, W2, 64), float("nan"), device=DEV)
    call("conv_tc_bn_x2_nhwc" if x2 else "conv_tc_bn_nhwc", ptr(xs), 0, 0, 0, N, H2, W2, 16, ptr(hi), ptr(lo), 4, ptr(db_), 64,
         0, 0, 0, 0, ptr(y), 0, ptr(ssum), ptr(ssq), ptr(dg), ptr(dbeta), 1e-5, 0.1, ptr(drm), ptr(drv), ptr(sc), ptr(sh),
         ptr(sm), ptr(si), ptr(ticket))
    torch.cuda.synchronize()
    err = relerr(from_nhwc(y), ref)
    print("tensor-core stem %s x2=%s: rel-to-max error %.2e" % (shape, x2, err))
    assert err < 2e-5
    assert relerr(ssum.cpu(), ref.sum(dim=(0, 2, 3))) < 2e-5
    assert relerr(ssq.cpu(), (ref * ref).sum(dim=(0, 2, 3))) < 2e-5
    mean = ref.mean(dim=(0, 2, 3))
    invstd = 1.0 / torch.sqrt(ref.var(dim=(0, 2, 3), unbiased=False) + 1e-5)
    assert relerr(sm.cpu(), mean) < 2e-5 and relerr(si.cpu(), invstd) < 2e-5
    assert int(ticket.item()) == 0
    # anything else with ksize = 4 is refused
    rc = lib().conv_tc_nhwc(ptr(xs), 0, 0, 0, N, H2, W2, 16, ptr(hi), 0, 4, ptr(db_), 64, 0, 0, 0, 0, ptr(y), 0, 0, 0, 0)
    assert rc != 0 and "k = 4" in lib().last_error()
