"""Hardware validation of the tcgen05-path building blocks against torch references."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fsnet_b200 import _lib, tc

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
OK = True


def check(name, got, ref, tol):
    global OK
    err = float((got.double() - ref.double()).norm() / (ref.double().norm() + 1e-30))
    good = err < tol and err == err
    OK &= good
    print(("OK   " if good else "FAIL ") + f"{name}: rel err {err:.2e} (tol {tol:.0e})", flush=True)


def planes_from(x, ring=1, c_pad=None):
    n, c, h, w = x.shape
    p = tc.Planes(n, h, w, c_pad or tc.pad16(c), ring, zero=True)
    _lib.call("fsnet_image_to_planes", x.contiguous(), c, p.view())
    return p


def test_conv_fwd_bwd(N, Cin, Cout, H, W, k, stride, pad, replicate):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
    xin = F.pad(x, (pad,) * 4, mode="replicate") if replicate else x
    y = F.conv2d(xin, w, stride=stride, padding=0 if replicate else pad)
    gy = torch.randn_like(y)
    gx_ref, gw_ref = torch.autograd.grad(y, (x, w), gy)
    tag = f"[{N},{Cin}->{Cout},{H}x{W},k{k},s{stride},p{pad},{'rep' if replicate else 'zero'}]"
    xp = planes_from(x.detach())
    check("planes roundtrip " + tag, xp.to_float()[:, :Cin], x.detach(), 1e-5)
    cw = tc.ConvWeights(w)
    cw.refresh(w)
    Ho, Wo = y.shape[-2:]
    out = tc.Fp32(N, Ho, Wo, cw.co_pad)
    stats = torch.zeros(2 * cw.co_pad, device="cuda", dtype=torch.float64)
    tc.conv(xp, cw, out, stride, pad, use_ring=replicate, stats=stats)
    check("conv fwd " + tag, out.nchw()[:, :Cout], y.detach(), 2e-5)
    check("conv stats " + tag, stats[:Cout], y.detach().double().sum((0, 2, 3)), 1e-4)
    # dy plane
    dy = tc.Planes(N, Ho, Wo, cw.co_pad, ring=0, zero=True)
    dy.t[0, :, :, :, :Cout] = gy.permute(0, 2, 3, 1).bfloat16()
    # weight gradient
    acc = tc.conv_wgrad(xp.view(), replicate, dy.view(), cw, stride, pad)
    gw = torch.zeros_like(w)
    _lib.call("fsnet_wgrad_to_param", acc, Cout, Cin, k, k, cw.co_pad, cw.ci_pad, gw, 0)
    check("conv wgrad " + tag, gw, gw_ref, 1e-2)
    # data gradient
    if stride == 1:
        if replicate:
            gx = tc.Fp32(N, H, W, cw.ci_pad, ring=1)
            full = tc.View(gx.t.data_ptr(), N, H + 2, W + 2, cw.ci_pad, 0, cw.ci_pad, 0)
            tc.conv_dgrad(dy, cw, full, pad=k - 1)
            _lib.call("fsnet_fold_ring", gx.view())
        else:
            gx = tc.Fp32(N, H, W, cw.ci_pad)
            tc.conv_dgrad(dy, cw, gx.view(), pad=k - 1 - pad)
    else:
        up = tc.Planes(N, H, W, cw.co_pad, ring=0)
        _lib.call("fsnet_zero_insert", dy.view(), up.view())
        gx = tc.Fp32(N, H, W, cw.ci_pad)
        tc.conv_dgrad(up, cw, gx.view(), pad=k - 1 - pad)
    check("conv dgrad " + tag, gx.nchw()[:, :Cin], gx_ref, 1e-2)


def test_bn_act(N, C, H, W, up, res_mode):
    g = torch.Generator(device="cuda").manual_seed(2)
    raw = torch.randn(N, C, H, W, device="cuda", generator=g).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    res = torch.randn(N, C, H, W, device="cuda", generator=g).requires_grad_(res_mode == 1)
    y = F.batch_norm(raw, rm.clone(), rv.clone(), gamma, beta, True, 0.1, 1e-5)
    if res_mode == 1:
        y = y + res
    a = F.relu(y)
    if up == 2:
        a_up = F.interpolate(a, scale_factor=2, mode="nearest")
    else:
        a_up = a
    ga = torch.randn_like(a_up)
    grads = torch.autograd.grad(a_up, (raw, gamma, beta) + ((res,) if res_mode == 1 else ()), ga)
    tag = f"[{N},{C},{H}x{W},up{up},res{res_mode}]"
    # ours
    rawb = tc.Fp32(N, H, W, C)
    rawb.t.copy_(raw.detach().permute(0, 2, 3, 1))
    stats = torch.cat([raw.detach().double().sum((0, 2, 3)), (raw.detach().double() ** 2).sum((0, 2, 3))]).contiguous()
    ss = torch.empty(2 * C, device="cuda"); mi = torch.empty(2 * C, device="cuda")
    nb = torch.zeros((), dtype=torch.long, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    cnt = N * H * W
    _lib.call("fsnet_bn_finalize", stats, tc.c_double(cnt), gamma.detach(), beta.detach(), None, rm2, rv2, nb, 0.1, 1e-5, 1, C, ss, mi)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    F.batch_norm(raw.detach(), rm_ref, rv_ref, gamma.detach(), beta.detach(), True, 0.1, 1e-5)
    check("bn running_mean " + tag, rm2, rm_ref, 1e-5)
    check("bn running_var " + tag, rv2, rv_ref, 1e-5)
    dst = tc.Planes(N, H * up, W * up, C + 16, ring=1, zero=True)
    resp = planes_from(res.detach()) if res_mode == 1 else None
    _lib.call("fsnet_act_planes", rawb.view(), ss, res_mode, resp.view() if resp else None, None, 1, up, dst.view(0, C))
    check("act planes " + tag, dst.to_float()[:, :C], a_up.detach(), 1e-5)
    # ring = replicate
    full = (dst.t[0].float() + dst.t[1].float())[..., :C].permute(0, 3, 1, 2)
    check("act ring " + tag, full, F.pad(a_up.detach(), (1, 1, 1, 1), mode="replicate"), 1e-5)
    # backward
    gbuf = tc.Fp32(N, H * up, W * up, C)
    gbuf.t.copy_(ga.permute(0, 2, 3, 1))
    act_lowres = tc.Planes(N, H, W, C, ring=1, zero=True)
    _lib.call("fsnet_act_planes", rawb.view(), ss, res_mode, resp.view() if resp else None, None, 1, 1, act_lowres.view())
    sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    _lib.call("fsnet_bn_bwd_reduce", gbuf.view(), up, act_lowres.view() if up == 1 else None, ss if up == 2 else None, rawb.view(), mi, sums)
    check("bn bwd dbeta " + tag, sums[:C], grads[2], 1e-4)
    check("bn bwd dgamma " + tag, sums[C:], grads[1], 1e-4)
    dy = tc.Planes(N, H, W, C, ring=0, zero=True)
    gres = tc.Fp32(N, H, W, C, zero=True)
    _lib.call("fsnet_bn_bwd_apply", gbuf.view(), up, act_lowres.view() if up == 1 else None, ss if up == 2 else None, rawb.view(), mi, gamma.detach(), sums, tc.c_double(cnt),
              dy.view(), 1 if res_mode == 1 else 0, gres.view() if res_mode == 1 else None, None, None, C)
    check("bn bwd dy " + tag, dy.t[0].float().permute(0, 3, 1, 2), grads[0], 1e-2)
    if res_mode == 1:
        check("bn bwd residual grad " + tag, gres.nchw(), grads[3], 1e-5)


def test_maxpool(N, C, H, W):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    x = (x.bfloat16().float()).requires_grad_(True)      # exactly representable: ties behave identically
    y = F.max_pool2d(x, 3, 2, 1)
    gy = torch.randn_like(y)
    gx_ref, = torch.autograd.grad(y, x, gy)
    xp = planes_from(x.detach())
    yp = tc.Planes(N, y.shape[2], y.shape[3], C, ring=1, zero=True)
    am = torch.empty(N, y.shape[2], y.shape[3], C, device="cuda", dtype=torch.uint8)
    _lib.call("fsnet_maxpool_planes", xp.view(), yp.view(), am)
    check(f"maxpool fwd [{N},{C},{H}x{W}]", yp.to_float(), y.detach(), 1e-6)
    gyb = tc.Fp32(N, y.shape[2], y.shape[3], C); gyb.t.copy_(gy.permute(0, 2, 3, 1))
    gxb = tc.Fp32(N, H, W, C, zero=True)
    _lib.call("fsnet_maxpool_bwd", xp.view(), am, gyb.view(), gxb.view(), 0)
    check(f"maxpool bwd [{N},{C},{H}x{W}]", gxb.nchw(), gx_ref, 1e-6)


def perf():
    """Forward / dgrad / wgrad times of the hot cfg2 layer shapes (B=12)."""
    shapes = [("(1,1) 96->32 @96x320 rep", 96, 32, 96, 320, 3, 1, 1, True), ("(0,1) 16->16 @192x640 rep", 16, 16, 192, 640, 3, 1, 1, True),
              ("disp1 32->16 @96x320 rep", 32, 16, 96, 320, 3, 1, 1, True), ("(0,0) 32->16 @96x320 zero", 32, 16, 96, 320, 3, 1, 1, False),
              ("layer1 64->64 @48x160", 64, 64, 48, 160, 3, 1, 1, False), ("layer4 512->512 @6x20", 512, 512, 6, 20, 3, 1, 1, False),
              ("stem 3->64 @192x640 s2", 3, 64, 192, 640, 7, 2, 3, False)]
    N = 12
    for name, Cin, Cout, H, W, k, stride, pad, rep in shapes:
        x = torch.randn(N, Cin, H, W, device="cuda")
        w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
        xp = planes_from(x)
        cw = tc.ConvWeights(w); cw.refresh(w)
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = tc.Fp32(N, Ho, Wo, cw.co_pad)
        stats = torch.zeros(2 * cw.co_pad, device="cuda", dtype=torch.float64)
        dy = tc.Planes(N, Ho, Wo, cw.co_pad, ring=0, zero=True)
        gx = tc.Fp32(N, H, W, cw.ci_pad, ring=1)
        full = tc.View(gx.t.data_ptr(), N, H + 2, W + 2, cw.ci_pad, 0, cw.ci_pad, 0)

        def t(fn, n=10):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e3 / n
        tf = t(lambda: tc.conv(xp, cw, out, stride, pad, use_ring=rep, stats=stats))
        tw = t(lambda: tc.conv_wgrad(xp.view(), rep, dy.view(), cw, stride, pad))
        td = float("nan")
        if stride == 1:
            td = t(lambda: tc.conv_dgrad(dy, cw, full if rep else gx.view(), pad=(k - 1) if rep else (k - 1 - pad)))
        print(f"PERF {name:32s} fwd {tf:7.1f} us  dgrad {td:7.1f} us  wgrad {tw:7.1f} us", flush=True)



if __name__ == "__main__":
    which = sys.argv[1:] or ["conv", "bn", "pool"]
    if "conv" in which:
        test_conv_fwd_bwd(2, 64, 64, 24, 40, 3, 1, 1, False)
        test_conv_fwd_bwd(2, 128, 256, 12, 20, 3, 1, 1, False)
        test_conv_fwd_bwd(2, 32, 16, 24, 40, 3, 1, 1, True)
        test_conv_fwd_bwd(2, 16, 16, 48, 80, 3, 1, 1, True)
        test_conv_fwd_bwd(2, 96, 32, 24, 40, 3, 1, 1, True)
        test_conv_fwd_bwd(2, 64, 128, 24, 40, 3, 2, 1, False)
        test_conv_fwd_bwd(2, 64, 128, 24, 40, 1, 2, 0, False)
        test_conv_fwd_bwd(2, 3, 64, 48, 80, 7, 2, 3, False)
        test_conv_fwd_bwd(2, 512, 512, 6, 10, 3, 1, 1, False)
        test_conv_fwd_bwd(2, 256, 12, 6, 10, 1, 1, 0, False)
    if "bn" in which:
        test_bn_act(2, 64, 12, 20, 1, 0)
        test_bn_act(2, 32, 12, 20, 2, 0)
        test_bn_act(3, 48, 7, 9, 1, 1)
    if "pool" in which:
        test_maxpool(2, 64, 24, 40)
        test_maxpool(2, 16, 13, 21)
    if "perf" in which:
        perf()
    print("ALL OK" if OK else "SOME FAILED")
    sys.exit(0 if OK else 1)


