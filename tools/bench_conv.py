"""CUDA-event timings of single convolution launches at cfg2a layer shapes (forward bf16x3 / data gradient), L2 flushed between launches.
FSNET_CONV_DBG / FSNET_CONV_OCC select diagnostic variants of the kernel (csrc/conv_tc.cu)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsnet_b200 import _lib, tc

LAYERS = {  # name: (N, Cin, Cout, H, W, k, stride, pad, replicate)
    "dec16_192x640": (12, 16, 16, 192, 640, 3, 1, 1, True),
    "dec32_96x320": (12, 32, 16, 96, 320, 3, 1, 1, True),
    "dec96_96x320": (12, 96, 32, 96, 320, 3, 1, 1, True),
    "l1_64_48x160": (12, 64, 64, 48, 160, 3, 1, 1, False),
    "l2_128_24x80": (12, 128, 128, 24, 80, 3, 1, 1, False),
    "l3_256_12x40": (12, 256, 256, 12, 40, 3, 1, 1, False),
    "l4_512_6x20": (12, 512, 512, 6, 20, 3, 1, 1, False),
}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


NOFLUSH = os.environ.get("BENCH_NOFLUSH", "0")


def timeit(fn, iters=10):
    ts = []
    for _ in range(iters):
        if NOFLUSH == "0":
            flush.zero_()
        elif NOFLUSH == "2":      # the same launch right before: warm L2, same shared-memory carve-out
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = {}
for name in (sys.argv[1:] or LAYERS):
    N, Cin, Cout, H, W, k, stride, pad, rep = LAYERS[name]
    w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
    cw = tc.ConvWeights(w); cw.refresh(w)
    xp = tc.Planes(N, H, W, cw.ci_pad, 1, zero=True)
    xp.t.normal_()
    out = tc.Fp32(N, H, W, cw.co_pad)
    stats = torch.zeros(2 * cw.co_pad, device="cuda", dtype=torch.float64)
    res[name + "/fwd"] = timeit(lambda: tc.conv(xp, cw, out, stride, pad, use_ring=rep, stats=stats))
    if rep:   # the executor's data gradient of a replicate-padded thin layer: dy with a zero ring of k-1, pad k-1, accumulate into the ringed gradient
        dy = tc.Planes(N, H, W, cw.co_pad, ring=k - 1, zero=True)
        dy.t[0].normal_()
        gx = tc.Fp32(N, H, W, cw.ci_pad, ring=1, zero=True)
        full = tc.View(gx.t.data_ptr(), N, H + 2, W + 2, cw.ci_pad, 0, cw.ci_pad, 0)
        res[name + "/dgrad_acc"] = timeit(lambda: tc.conv_dgrad(dy, cw, full, pad=k - 1, accumulate=True, use_ring=True))
        res[name + "/dgrad"] = timeit(lambda: tc.conv_dgrad(dy, cw, full, pad=k - 1, accumulate=False, use_ring=True))
    else:
        dy = tc.Planes(N, H, W, cw.co_pad, ring=0, zero=True)
        dy.t[0].normal_()
        gx = tc.Fp32(N, H, W, cw.ci_pad)
        res[name + "/dgrad"] = timeit(lambda: tc.conv_dgrad(dy, cw, gx.view(), pad=k - 1 - pad))
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
