"""Micro-benchmark of the fused loss kernels at cfg2 size (B=12, 192x640): per-scale forward/backward
time with CUDA events, L2 flushed between launches; prints achieved algorithmic GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fsnet_b200 import _lib
from fsnet_b200.data.synthetic import make_batch


def main(B=12, H=192, W=640, iters=20):
    dev = "cuda"
    data = make_batch(B, H, W)
    g = torch.Generator().manual_seed(5)
    outs = {}
    for s in range(4):       # smooth random depth pyramid, 2..40 m
        h, w = H >> s, W >> s
        f = torch.nn.functional.interpolate(torch.rand(B, 1, max(h // 8, 2), max(w // 8, 2), generator=g), size=(h, w), mode="bilinear", align_corners=True)
        outs[("depth", s, s)] = (2.0 * torch.exp(f * 3.0)).contiguous()
    tgt, s0, s1 = (data[("original_image", f)].to(dev) for f in (0, 1, -1))
    mask = data["patched_mask"].float().to(dev)
    cam = torch.empty(B, 2, 21, device=dev)
    _lib.call("fsnet_camera_setup", data["P2"].to(dev), data[("relative_pose", 1)].to(dev), data[("relative_pose", -1)].to(dev), B, cam)
    ident = torch.empty(B, 2, H, W, device=dev)
    packed = torch.empty(3, B, H, W, 4, device=dev)
    noise = torch.randn(B, 2, H, W, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    gout = torch.full((1,), 0.25, device=dev)
    res = {}

    def timeit(fn):
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    res["identity_us"] = timeit(lambda: _lib.call("fsnet_identity_photometric_masked", tgt, s0, s1, mask, 1, B, H, W, ident, packed))
    for s in range(4):
        d = outs[("depth", s, s)].to(dev)
        hs, ws = d.shape[-2:]
        acc = torch.zeros(4, dtype=torch.float64, device=dev)
        gd = torch.zeros_like(d)
        f = lambda: _lib.call("fsnet_warp_ssim_fwd", d, hs, ws, packed, mask, 1, cam, ident, noise, None,
                              _lib.ctypes.c_uint(1), B, H, W, acc, None, None)
        bwd = lambda: _lib.call("fsnet_warp_ssim_bwd", d, hs, ws, packed, mask, 1, cam, ident, noise, None,
                                _lib.ctypes.c_uint(1), B, H, W, acc, gout, gd, None)
        acc2 = torch.zeros(4, dtype=torch.float64, device=dev)
        acc2[1] = float(mask.sum())
        unit = torch.full((1,), 0.25, device=dev)

        def fused():
            gd.zero_()
            _lib.call("fsnet_warp_ssim_fwdbwd", None, None, d, hs, ws, packed, mask, 1, cam, ident, noise, None,
                      _lib.ctypes.c_uint(5), B, H, W, acc2, unit, gd, None)
        fused_only = lambda: _lib.call("fsnet_warp_ssim_fwdbwd", None, None, d, hs, ws, packed, mask, 1, cam, ident, noise, None,
                                       _lib.ctypes.c_uint(5), B, H, W, acc2, unit, gd, None)
        tf, tb = timeit(f), timeit(bwd)
        tfu = timeit(fused_only)
        res[f"fused_s{s}_us"] = tfu
        res[f"fused_s{s}_GBs"] = B * H * W * (40 + 16 / 4 ** s) / tfu / 1e3
        bytes_f = B * H * W * (40 + 8 / 4 ** s)
        bytes_b = B * H * W * (40 + 16 / 4 ** s)
        res[f"fwd_s{s}_us"] = tf
        res[f"fwd_s{s}_GBs"] = bytes_f / tf / 1e3
        res[f"bwd_s{s}_us"] = tb
        res[f"bwd_s{s}_GBs"] = bytes_b / tb / 1e3
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
