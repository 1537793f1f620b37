"""Hardware validation of fsnet_conv_fwd (tcgen05 implicit GEMM) against F.conv2d in fp64."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fsnet_b200 import _lib


def to_planes(x, ring, replicate):
    """fp32 NCHW -> (hi, lo) bf16 NHWC with a `ring`-pixel border (replicate or zeros)."""
    if ring:
        x = F.pad(x, (ring, ring, ring, ring), mode="replicate" if replicate else "constant")
    x = x.permute(0, 2, 3, 1).contiguous()
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.contiguous(), lo.contiguous()


def w_planes(w):
    w = w.permute(0, 2, 3, 1).contiguous()     # [Cout, KH, KW, Cin]
    hi = w.bfloat16()
    lo = (w - hi.float()).bfloat16()
    return hi.contiguous(), lo.contiguous()


def run(N, Cin, Cout, H, W, k, stride, pad, replicate=False, nprod=3, bias=False, relu=False, stats=True, seed=0, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) if bias else None
    ring = 1 if (replicate or True) else 0
    ring = max(ring, pad if replicate else 1)
    hi, lo = to_planes(x, ring, replicate)
    whi, wlo = w_planes(w)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.full((N, Ho, Wo, Cout), float("nan"), device="cuda")
    st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64) if stats else None
    args = (hi, lo, N, H, W, Cin, W + 2 * ring, H + 2 * ring, ring, whi, wlo, Cout, k, k, stride, pad, int(replicate), nprod, b, int(relu), out, st)
    _lib.call("fsnet_conv_fwd", *args)
    torch.cuda.synchronize()
    xr = x.double()
    if replicate:
        ref = F.conv2d(F.pad(xr, (pad, pad, pad, pad), mode="replicate"), w.double(), None if b is None else b.double(), stride=stride)
    else:
        ref = F.conv2d(xr, w.double(), None if b is None else b.double(), stride=stride, padding=pad)
    if relu:
        ref = ref.relu()
    got = out.permute(0, 3, 1, 2).double()
    err = float((got - ref).norm() / ref.norm())
    msg = f"N={N} Cin={Cin} Cout={Cout} {H}x{W} k={k} s={stride} p={pad} rep={int(replicate)} nprod={nprod}: rel err {err:.2e}"
    if stats:
        s1 = ref.sum((0, 2, 3)); s2 = (ref * ref).sum((0, 2, 3))
        e1 = float((st[:Cout] - s1).abs().max() / (s1.abs().max() + 1e-9)); e2 = float((st[Cout:] - s2).abs().max() / s2.abs().max())
        msg += f" stats err {e1:.1e} {e2:.1e}"
    if time_it:
        for _ in range(3):
            _lib.call("fsnet_conv_fwd", *args)
        torch.cuda.synchronize()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            _lib.call("fsnet_conv_fwd", *args)
        bb.record(); torch.cuda.synchronize()
        us = a.elapsed_time(bb) * 100
        fl = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        msg += f" | {us:.1f} us, {fl / us / 1e6:.1f} TFLOP/s useful ({nprod}x MMA work)"
    tol = 2e-5 if nprod == 3 else 1e-2
    print(("OK   " if (err < tol and err == err) else "FAIL ") + msg, flush=True)
    return err < tol


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "basic"
    ok = True
    if which == "basic":
        ok &= run(1, 64, 64, 8, 16, 3, 1, 1)                       # one tile, SW128
        ok &= run(2, 64, 64, 48, 160, 3, 1, 1, time_it=True)       # layer1 shape
        ok &= run(2, 128, 128, 24, 80, 3, 1, 1)
        ok &= run(2, 256, 256, 12, 40, 3, 1, 1)                    # 2 n-tiles
        ok &= run(2, 512, 512, 6, 20, 3, 1, 1, time_it=True)
        ok &= run(2, 32, 16, 96, 320, 3, 1, 1, replicate=True)     # SW64, replicate ring
        ok &= run(2, 16, 16, 192, 640, 3, 1, 1, replicate=True, bias=True, stats=False)   # SW32, dispconv-like
        ok &= run(2, 96, 32, 96, 320, 3, 1, 1, replicate=True)
        ok &= run(2, 64, 64, 13, 26, 1, 1, 0)                      # 1x1
        ok &= run(2, 64, 64, 48, 160, 3, 1, 1, nprod=1)
    elif which == "stride":
        ok &= run(2, 64, 128, 48, 160, 3, 2, 1)
        ok &= run(2, 64, 128, 48, 160, 1, 2, 0)
        ok &= run(2, 16, 64, 192, 640, 7, 2, 3)                    # stem with padded channels
        ok &= run(2, 64, 128, 47, 159, 3, 2, 1)                    # odd sizes
    elif which == "perf":
        for shp in [(12, 64, 64, 48, 160, 3, 1, 1), (12, 128, 128, 24, 80, 3, 1, 1), (12, 256, 256, 12, 40, 3, 1, 1),
                    (12, 512, 512, 6, 20, 3, 1, 1), (12, 16, 16, 192, 640, 3, 1, 1), (12, 96, 32, 96, 320, 3, 1, 1),
                    (12, 512, 256, 12, 40, 3, 1, 1), (12, 16, 64, 192, 640, 7, 2, 3)]:
            ok &= run(*shp, time_it=True)
    sys.exit(0 if ok else 1)
