"""Thin Python handles for the tcgen05 path: ringed NHWC buffers and the calls of include/fsnet_b200.h.

``Planes``  bf16 [2, N, H+2r, W+2r, C]  (hi, lo) -- what the convolutions read through TMA
``Fp32``    fp32 [N, H+2r, W+2r, C]             -- conv outputs ("raw") and gradients
Both hand out ``fsnet_view`` structs (optionally a channel slice).  Product code: CUDA only.
"""
import ctypes

import torch

from . import _lib
from ._lib import View


def pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def pad_in(c: int) -> int:
    """Input-channel padding of a convolution: multiples of 16, except the network stems (3 / 6 image channels -> 8: the seven
    x-taps of a 7x7 kernel row are then 56 contiguous elements = one 64-element K slice of the folded-tap path, csrc/conv_tc.cu)."""
    return 8 if c <= 8 else pad16(c)


class _Buf:
    def __init__(self, t, n, h, w, c, ring):
        self.t, self.n, self.h, self.w, self.c, self.ring = t, n, h, w, c, ring

    def view(self, c_off=0, c=None) -> View:
        return View(self.t.data_ptr(), self.n, self.h, self.w, self.c if c is None else c, self.ring, self.c, c_off)


class Planes(_Buf):
    SLACK = 256      # zeroed elements behind the lo plane: the folded-tap TMA windows of the last pixels run past the end

    def __init__(self, n, h, w, c, ring=1, device="cuda", zero=False, slack=None):
        numel = 2 * n * (h + 2 * ring) * (w + 2 * ring) * c
        if slack is None:            # only the folded-tap convolution path (thin layers, 96-channel concat) over-reads
            slack = c < 64 or c == 96
        flat = (torch.zeros if zero else torch.empty)(numel + (self.SLACK if slack else 0), device=device, dtype=torch.bfloat16)
        if slack and not zero:
            flat[numel:].zero_()
        self._flat = flat
        t = flat[:numel].view(2, n, h + 2 * ring, w + 2 * ring, c)
        super().__init__(t, n, h, w, c, ring)

    def to_float(self):
        """[N,C,H,W] fp32 reconstruction of the interior (tests / lazy feature export)."""
        r = self.ring
        x = self.t[0].float() + self.t[1].float()
        return x[:, r:r + self.h, r:r + self.w, :].permute(0, 3, 1, 2)


class Fp32(_Buf):
    def __init__(self, n, h, w, c, ring=0, device="cuda", zero=False):
        t = (torch.zeros if zero else torch.empty)(n, h + 2 * ring, w + 2 * ring, c, device=device, dtype=torch.float32)
        super().__init__(t, n, h, w, c, ring)

    @classmethod
    def wrap(cls, t: torch.Tensor):
        """View an existing contiguous fp32 [N,H,W,C] tensor (no ring) as a buffer."""
        obj = cls.__new__(cls)
        _Buf.__init__(obj, t, t.shape[0], t.shape[1], t.shape[2], t.shape[3], 0)
        return obj

    def interior(self):
        r = self.ring
        return self.t[:, r:r + self.h, r:r + self.w, :]

    def nchw(self):
        return self.interior().permute(0, 3, 1, 2)


class ConvWeights:
    """bf16 operand planes of one convolution, refreshed from the fp32 parameter every step."""

    def __init__(self, weight: torch.Tensor, need_dgrad=True, ci_pad=None):
        co, ci, kh, kw = weight.shape
        self.co, self.ci, self.kh, self.kw = co, ci, kh, kw
        self.co_pad, self.ci_pad = pad16(co), (pad16(ci) if ci_pad is None else ci_pad)      # ci_pad = 8: the network stem (pad_in)
        dev = weight.device
        self.fwd = torch.empty(2, self.co_pad, kh, kw, self.ci_pad, device=dev, dtype=torch.bfloat16)
        self.dgrad = torch.empty(self.ci_pad, kh, kw, self.co_pad, device=dev, dtype=torch.bfloat16) if need_dgrad else None
        self.acc = None

    def refresh(self, weight: torch.Tensor):
        _lib.call("fsnet_weight_planes", weight.detach(), self.co, self.ci, self.kh, self.kw, self.co_pad, self.ci_pad,
                  self.fwd[0], self.fwd[1], self.dgrad)

    @property
    def acc_numel(self):
        return self.co_pad * self.kh * self.kw * self.ci_pad

    def desc(self, weight):
        d = _lib.WeightDesc()
        d.w = weight.data_ptr(); d.fwd_hi = self.fwd[0].data_ptr(); d.fwd_lo = self.fwd[1].data_ptr()
        d.dgrad_hi = 0 if self.dgrad is None else self.dgrad.data_ptr()
        d.cout, d.cin, d.kh, d.kw, d.cout_pad, d.cin_pad = self.co, self.ci, self.kh, self.kw, self.co_pad, self.ci_pad
        return d


def conv(inp: Planes, w: ConvWeights, out: Fp32, stride=1, pad=1, use_ring=False, nprod=3, bias=None, relu=False, stats=None,
         accumulate=False, in_view=None, out_view=None):
    _lib.call("fsnet_conv", in_view or inp.view(), int(use_ring), w.fwd[0], w.fwd[1], w.co_pad, w.kh, w.kw, stride, pad, nprod,
              bias, int(relu), out_view or out.view(), int(accumulate), stats)


def conv_dgrad(dy: Planes, w: ConvWeights, out_view: View, pad, accumulate=False, dy_view=None, use_ring=False):
    """Data gradient of a stride-1 convolution = convolution of dy with the flipped, transposed weights.
    ``use_ring``: dy carries a materialised zero ring of width ``pad`` (enables the folded-tap path for thin layers)."""
    _lib.call("fsnet_conv", dy_view or dy.view(), int(use_ring), w.dgrad, None, w.ci_pad, w.kh, w.kw, 1, pad, 1, None, 0, out_view,
              int(accumulate), None)


def conv_wgrad(x_view: View, use_ring, dy_view: View, w: ConvWeights, stride, pad, acc=None):
    """acc: zeroed fp32 [co_pad, kh, kw, ci_pad] accumulator (a slice of the tape's pool) or None (allocated here)."""
    if acc is None:
        acc = torch.zeros(w.acc_numel, device=w.fwd.device, dtype=torch.float32)
    _lib.call("fsnet_conv_wgrad", x_view, int(use_ring), dy_view, w.kh, w.kw, stride, pad, acc)
    return acc


def c_double(x):
    return ctypes.c_double(float(x))
