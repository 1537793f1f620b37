"""tcgen05 convolution kernels (forward bf16x3, data gradient, weight gradient) against torch.nn.functional.conv2d in fp32,
layer shapes of the in-scope networks incl. the folded-tap paths (thin replicate layers, the 7x7/2 stem on a zero ring)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


CASES = [
    # N, Cin, Cout, H, W, k, stride, pad, mode      mode: "zero" (TMA OOB fill), "rep" (replicate ring), "zring" (materialised zero ring)
    (2, 64, 64, 24, 40, 3, 1, 1, "zero"),
    (2, 16, 16, 48, 80, 3, 1, 1, "rep"),
    (3, 16, 16, 38, 70, 3, 1, 1, "rep"),        # partial tiles in both directions
    (2, 32, 16, 24, 48, 3, 1, 1, "rep"),
    (2, 96, 32, 24, 48, 3, 1, 1, "rep"),
    (2, 3, 64, 48, 80, 7, 2, 3, "zring"),
    (3, 6, 64, 64, 96, 7, 2, 3, "zring"),
    (2, 3, 64, 48, 80, 7, 2, 3, "zring8"),      # the networks' stem as the executor runs it: image planes padded to 8 channels
    (3, 6, 64, 64, 96, 7, 2, 3, "zring8"),      # (7 taps x 8 channels = one 64-element K slice per kernel row)
    (2, 3, 64, 192, 640, 7, 2, 3, "zring8"),    # full cfg2 image: 2-row x 64-pixel tiles, K-split weight gradient
    (2, 3, 64, 48, 80, 7, 2, 3, "zero"),
    (2, 64, 128, 24, 40, 3, 2, 1, "zero"),
    (2, 256, 512, 6, 10, 3, 1, 1, "zero"),
    (2, 128, 256, 12, 20, 1, 2, 0, "zero"),
    # halo path (3x3 / stride 1, 64-channel chunks): 16 x 8 and 8 x 16 pixel tiles, several chunks, two channel tiles, partial tiles, ring
    (2, 64, 64, 48, 40, 3, 1, 1, "zero"),
    (2, 128, 128, 24, 80, 3, 1, 1, "zero"),
    (2, 256, 256, 12, 40, 3, 1, 1, "zero"),
    (2, 128, 64, 20, 36, 3, 1, 1, "rep"),
    (3, 64, 32, 50, 70, 3, 1, 1, "rep"),
]


@pytest.mark.parametrize("N,Cin,Cout,H,W,k,stride,pad,mode", CASES)
def test_conv_forward_and_gradients(N, Cin, Cout, H, W, k, stride, pad, mode):
    run_conv_case("cuda", N, Cin, Cout, H, W, k, stride, pad, mode)


def run_conv_case(dev, N, Cin, Cout, H, W, k, stride, pad, mode):
    """``dev`` = "cuda" here; tests/test_emulated_kernels_cpu.py runs the same checks on "cpu" against the ABI-level stand-ins."""
    from fsnet_b200 import _lib, tc
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(N, Cin, H, W, device=dev, generator=g).requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, device=dev, generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
    rep = mode == "rep"
    xin = F.pad(x, (pad,) * 4, mode="replicate") if rep else x
    y = F.conv2d(xin, w, stride=stride, padding=0 if rep else pad)
    gy = torch.randn(y.shape, device=dev, generator=g)
    gx_ref, gw_ref = torch.autograd.grad(y, (x, w), gy)
    ring = pad if mode.startswith("zring") else 1
    ci_pad = tc.pad_in(Cin) if mode == "zring8" else tc.pad16(Cin)
    xp = tc.Planes(N, H, W, ci_pad, ring, zero=True, device=dev)
    _lib.call("fsnet_image_to_planes_ring", x.detach().contiguous(), Cin, xp.view(), int(mode.startswith("zring")))
    use_ring = mode in ("rep", "zring", "zring8")
    cw = tc.ConvWeights(w, ci_pad=ci_pad)
    cw.refresh(w)
    Ho, Wo = y.shape[-2:]
    out = tc.Fp32(N, Ho, Wo, cw.co_pad, device=dev)
    stats = torch.zeros(2 * cw.co_pad, device=dev, dtype=torch.float64)
    tc.conv(xp, cw, out, stride, pad, use_ring=use_ring, stats=stats)
    assert rel(out.nchw()[:, :Cout], y.detach()) < 2e-5                      # bf16x3 split operands: fp32-class accuracy
    assert rel(stats[:Cout], y.detach().double().sum((0, 2, 3))) < 1e-4       # fused BatchNorm statistics
    assert rel(stats[cw.co_pad:cw.co_pad + Cout], (y.detach().double() ** 2).sum((0, 2, 3))) < 1e-4
    # weight gradient (single bf16 product)
    dy = tc.Planes(N, Ho, Wo, cw.co_pad, ring=0, zero=True, device=dev)
    dy.t[0, :, :, :, :Cout] = gy.permute(0, 2, 3, 1).bfloat16()
    acc = tc.conv_wgrad(xp.view(), use_ring, dy.view(), cw, stride, pad)
    gw = torch.zeros_like(w)
    _lib.call("fsnet_wgrad_to_param", acc, Cout, Cin, k, k, cw.co_pad, cw.ci_pad, gw, 0)
    assert rel(gw, gw_ref) < 1e-2, rel(gw, gw_ref)
    # per-tap check: a permutation of taps would keep the norm but not this
    assert rel(gw[:, :, 0, k - 1], gw_ref[:, :, 0, k - 1]) < 2e-2 and rel(gw[:, :, k - 1, 0], gw_ref[:, :, k - 1, 0]) < 2e-2
    # data gradient of the stride-1 layers
    if stride == 1:
        if rep:
            gx = tc.Fp32(N, H, W, cw.ci_pad, ring=1, device=dev)
            full = tc.View(gx.t.data_ptr(), N, H + 2, W + 2, cw.ci_pad, 0, cw.ci_pad, 0)
            tc.conv_dgrad(dy, cw, full, pad=k - 1)
            _lib.call("fsnet_fold_ring", gx.view())
            # the executor's variant: dy with a materialised zero ring of k-1 pixels read as data (folded-tap / TMEM-operand paths)
            dyr = tc.Planes(N, Ho, Wo, cw.co_pad, ring=k - 1, zero=True, device=dev)
            dyr.t[0, :, k - 1:k - 1 + Ho, k - 1:k - 1 + Wo, :Cout] = gy.permute(0, 2, 3, 1).bfloat16()
            gx2 = tc.Fp32(N, H, W, cw.ci_pad, ring=1, device=dev)
            full2 = tc.View(gx2.t.data_ptr(), N, H + 2, W + 2, cw.ci_pad, 0, cw.ci_pad, 0)
            tc.conv_dgrad(dyr, cw, full2, pad=k - 1, use_ring=True)
            _lib.call("fsnet_fold_ring", gx2.view())
            assert rel(gx2.nchw()[:, :Cin], gx_ref) < 1e-2
        else:
            gx = tc.Fp32(N, H, W, cw.ci_pad, device=dev)
            tc.conv_dgrad(dy, cw, gx.view(), pad=k - 1 - pad)
        assert rel(gx.nchw()[:, :Cin], gx_ref) < 1e-2
