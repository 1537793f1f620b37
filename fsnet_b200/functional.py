"""torch.autograd Functions over the fsnet_b200 C ABI (include/fsnet_b200.h).

Every Function here launches hand-written CUDA kernels through ctypes; none has a CPU or PyTorch
fallback.  Reference citations (file:line) are relative to the FSNet checkout.
"""
from typing import List, Optional, Sequence

import os

import torch

from . import _lib

MASK_NONE, MASK_F32, MASK_F64 = 0, 1, 2
FLAG_OVERLAP, FLAG_MOTION, FLAG_PACKED_MASK = 1, 2, 4


def _mask_dtype(mask: Optional[torch.Tensor]) -> int:
    if mask is None:
        return MASK_NONE
    if mask.dtype == torch.float64:
        return MASK_F64
    if mask.dtype == torch.float32:
        return MASK_F32
    raise _lib.FsnetError(f"patched_mask must be float32 or float64, got {mask.dtype}")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _prep_mask(mask):
    mask_c = None if mask is None else mask.detach()
    if mask_c is not None and not mask_c.is_floating_point():
        mask_c = mask_c.float()           # e.g. the uint8 fisheye validity image: the reference's products promote it to fp32
    if mask_c is not None:
        mask_c = mask_c.contiguous()
    return mask_c, _mask_dtype(mask_c)


def _fused_scales(scales, log_image, need_grad=True):
    # per scale: only the scale that feeds the log images (scale 0 of a log-image head) keeps separate launches
    return [need_grad and not (log_image and s == 0) for s in scales]


def _batch_only_terms(tgt, src0, src1, mask_c, mdt, scales, disp_hw, fused_s):
    """The part of the loss that depends on the BATCH only, not on the networks' outputs: identity photometric terms + the
    RGBX-packed frames, the normaliser sum(patched_mask) in the accumulator rows of the fused scales, the colour pyramid of the
    smoothness term.  Runs where it is called (the loss itself, or ``prefetch_batch_terms`` on a side stream at step start)."""
    dev = tgt.device
    B, _, H, W = tgt.shape
    S = len(scales)
    ident = torch.empty(B, 2, H, W, device=dev, dtype=torch.float32)
    packed = torch.empty(3, B, H, W, 4, device=dev, dtype=torch.float32)
    # (the packed pixels' 4th component carries patched_mask: the frame-pair kernel reads loss weight and overlap mask from it)
    _lib.call("fsnet_identity_photometric_masked", tgt, src0, src1, mask_c, mdt, B, H, W, ident, packed)
    acc = torch.zeros(S, 4, device=dev, dtype=torch.float64)
    if any(fused_s):
        # the normaliser sum(patched_mask) goes ONLY into the rows of fused scales: the forward-only kernel of a non-fused
        # scale (scale 0 of a log-image head) accumulates its own into acc[i, 1]
        i0 = fused_s.index(True)
        if all(fused_s[i0:]):
            _lib.call("fsnet_mask_sum", mask_c, mdt, _lib.ctypes.c_longlong(B * H * W), acc[i0:], S - i0, 4)
        else:
            for i in range(S):
                if fused_s[i]:
                    _lib.call("fsnet_mask_sum", mask_c, mdt, _lib.ctypes.c_longlong(B * H * W), acc[i], 1, 4)
    colour = []
    for (h, w) in disp_hw:
        if h == H and w == W:
            colour.append(tgt)
        else:                      # colour image of the smoothness term at this scale (adaptive_avg_pool2d), shared with backward
            pooled = torch.empty(B, 3, h, w, device=dev, dtype=torch.float32)
            _lib.call("fsnet_box_pool", tgt, B * 3, H, W, H // h, pooled)
            colour.append(pooled)
    return ident, packed, acc, colour


class BatchTerms:
    """Result of ``prefetch_batch_terms``: tensors produced on a side stream + the event the consumer waits for."""
    __slots__ = ("key", "event", "ident", "packed", "acc", "colour", "noise", "mask_c", "mdt", "fused_s")


_SIDE = {}


def prefetch_batch_terms(tgt, src0, src1, mask, scales, log_image, draw_noise: bool):
    """Launch the batch-only part of the loss (see ``_batch_only_terms``; ~120 us at cfg2a: identity terms 43, mask sum 18, colour
    pyramid 28, tie-break noise 27) on a side stream at the START of the step, under the encoder, instead of between the decoder
    and the fused loss kernels.  Returns a BatchTerms for ``reprojection_loss(..., pre=...)``; CUDA only, training steps only."""
    dev = tgt.device
    side = _SIDE.get(dev)
    if side is None:
        side = _SIDE[dev] = torch.cuda.Stream(device=dev)
    B, _, H, W = tgt.shape
    pre = BatchTerms()
    pre.key = (tgt.data_ptr(), src0.data_ptr(), src1.data_ptr(), None if mask is None else mask.data_ptr(), tuple(scales), bool(log_image))
    side.wait_stream(torch.cuda.current_stream())                 # the batch is on the device
    with torch.cuda.stream(side):
        t, a, b = _f32c(tgt), _f32c(src0), _f32c(src1)
        pre.mask_c, pre.mdt = _prep_mask(mask)
        pre.fused_s = _fused_scales(scales, log_image)
        disp_hw = [(H >> s, W >> s) for s in scales]
        pre.ident, pre.packed, pre.acc, pre.colour = _batch_only_terms(t, a, b, pre.mask_c, pre.mdt, scales, disp_hw, pre.fused_s)
        pre.noise = list(torch.randn(len(scales), B, 2, H, W, device=dev).unbind(0)) if draw_noise else None
        pre.event = torch.cuda.Event()
        pre.event.record()
    return pre


class _ReprojectionLoss(torch.autograd.Function):
    """MonoDepth2Decoder.compute_total_reprojection_loss (monodepth2_decoder.py:205-304) as
    camera set-up + identity terms + one fused kernel per scale + smoothness + finalise.

    Inputs: S depth maps, S disparity maps, the two cam_T_cam matrices, then constants.
    Output: (total [0-d], stats [2S+2] fp64 = loss/s, smooth_loss/s, total, 0, sel, pred0) with
    sel [B,H,W] uint8 = arg-min index at scale 0 and pred0 [2,3,H,W] = warped sources of sample 0
    (both empty unless cfg['log_image']; they feed the reference's `hm` dict, :223,237-238,265-268).
    """

    @staticmethod
    def forward(ctx, S: int, cfg: dict, *tensors):
        depths = [_f32c(t) for t in tensors[:S]]
        disps = [_f32c(t) for t in tensors[S:2 * S]]
        T0, T1, P2, tgt, src0, src1, mask, motion = tensors[2 * S:2 * S + 8]
        noise = tensors[2 * S + 8:]
        dev = tgt.device
        B, _, H, W = tgt.shape
        tgt, src0, src1 = _f32c(tgt), _f32c(src0), _f32c(src1)
        P2c, T0c, T1c = _f32c(P2), _f32c(T0), _f32c(T1)
        mask_c, mdt = _prep_mask(mask)
        flags = (FLAG_OVERLAP if cfg["overlapped_mask"] else 0) | (FLAG_MOTION if motion is not None else 0)
        motion_c = None if motion is None else _f32c(motion)
        noise_c = [None] * S if len(noise) == 0 else [_f32c(n) for n in noise]

        cam = torch.empty(B, 2, 21, device=dev, dtype=torch.float32)
        mei = cfg.get("mei")            # MEI fisheye camera: dict(calib [B,3], lut [B,H,W,4], lut_idx [B]) from mei_ray_table
        if mei is None:
            _lib.call("fsnet_camera_setup", P2c, T0c, T1c, B, cam)
            warp_fwd, lut_args = "fsnet_warp_ssim_fwd", ()
        else:
            _lib.call("fsnet_camera_setup_mei", P2c, mei["calib"], T0c, T1c, B, cam)
            warp_fwd, lut_args = "fsnet_warp_ssim_mei_fwd", (mei["lut"], mei["lut_idx"])
        # batch-only terms: identity photometric map + RGBX-packed frames, mask sums, colour pyramid -- prefetched on a side stream at
        # step start when the head asked for it (prefetch_batch_terms), else computed here
        need_grad = any(ctx.needs_input_grad[2 + i] for i in range(S)) or any(ctx.needs_input_grad[2 + 2 * S + j] for j in range(2))
        fused_s = _fused_scales(cfg["scales"], bool(cfg.get("log_image")), need_grad)
        disp_hw = [tuple(d.shape[-2:]) for d in disps]
        pre = cfg.get("pre")
        key = (tgt.data_ptr(), src0.data_ptr(), src1.data_ptr(), None if mask is None else mask.data_ptr(), tuple(cfg["scales"]), bool(cfg.get("log_image")))
        if pre is not None:
            torch.cuda.current_stream().wait_event(pre.event)        # always joined, also when the result turns out not to fit
        if pre is not None and pre.key == key and pre.fused_s == fused_s and [tuple(c.shape[-2:]) for c in pre.colour] == disp_hw:
            main = torch.cuda.current_stream()
            ident, packed, acc, colour = pre.ident, pre.packed, pre.acc, pre.colour
            for t_ in [ident, packed, acc] + [c for c in colour if c is not tgt]:
                t_.record_stream(main)
        else:
            ident, packed, acc, colour = _batch_only_terms(tgt, src0, src1, mask_c, mdt, cfg["scales"], disp_hw, fused_s)
        flags |= FLAG_PACKED_MASK
        if motion is not None:
            ident = None
        sums = torch.zeros(S, B, 3, device=dev, dtype=torch.float64)
        sel = pred0 = None
        if cfg.get("log_image"):
            sel = torch.empty(B, H, W, device=dev, dtype=torch.uint8)
            pred0 = torch.empty(2, 3, H, W, device=dev, dtype=torch.float32)
        ctx.need_pose = any(ctx.needs_input_grad[2 + 2 * S + j] for j in range(2))
        # Training steps: ONE launch per scale computes the forward sums AND d loss / d depth (for the unit upstream gradient
        # 1/S; backward() rescales -- the loss is linear in it).  The log-image outputs need the forward-only kernel.
        fused = any(fused_s)
        ctx.fused = fused_s
        unit_gd, unit_gP = [None] * S, None
        if fused:
            unit = torch.full((1,), 1.0 / S, device=dev, dtype=torch.float32)
            unit_gP = torch.zeros(B, 2, 12, device=dev, dtype=torch.float32) if ctx.need_pose else None
        # Smoothness term on a side stream, under the frame-pair kernels (which are issue bound and leave the memory system idle):
        # forward sums and -- in training steps -- the backward for the unit upstream gradient, which backward() rescales like
        # the fused photometric gradients.  Joined before the loss is finalised.
        side = _SIDE.get(dev) if (dev.type == "cuda" and os.environ.get("FSNET_SMOOTH_STREAM", "1") != "0") else None
        if side is None and dev.type == "cuda" and os.environ.get("FSNET_SMOOTH_STREAM", "1") != "0":
            side = _SIDE[dev] = torch.cuda.Stream(device=dev)
        unit_gs = [None] * S
        unit_smooth = need_grad and side is not None

        def smooth_terms():
            for i, s in enumerate(cfg["scales"]):
                h, w = disps[i].shape[-2:]
                wgt = float(cfg["smooth_weight"] / (2 ** s))
                _lib.call("fsnet_smooth_fwd", disps[i], colour[i], B, h, w, h, w, wgt, sums[i], acc[i, 2:])
                if unit_smooth:
                    unit_gs[i] = torch.empty(disps[i].shape, device=dev, dtype=torch.float32)
                    _lib.call("fsnet_smooth_bwd", disps[i], colour[i], B, h, w, h, w, wgt, sums[i], unit_s, unit_gs[i])

        if side is not None:
            unit_s = torch.full((1,), 1.0 / S, device=dev, dtype=torch.float32)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                smooth_terms()
        for i, s in enumerate(cfg["scales"]):
            hs, ws = depths[i].shape[-2:]
            want_log = sel is not None and s == 0
            if fused_s[i]:
                gd = torch.zeros(depths[i].shape, device=dev, dtype=torch.float32)     # the fused kernel ADDS its partial gradients
                _lib.call("fsnet_warp_ssim_fwdbwd", *(lut_args if lut_args else (None, None)), depths[i], hs, ws, packed, mask_c, mdt, cam,
                          ident, noise_c[i], motion_c, _lib.ctypes.c_uint(flags), B, H, W, acc[i], unit, gd, unit_gP)
                unit_gd[i] = gd
            else:
                _lib.call(warp_fwd, *lut_args, depths[i], hs, ws, packed, mask_c, mdt, cam, ident, noise_c[i],
                          motion_c, _lib.ctypes.c_uint(flags), B, H, W, acc[i], sel if want_log else None,
                          pred0 if want_log else None)
        if side is not None:
            main = torch.cuda.current_stream()
            main.wait_stream(side)
            for t_ in unit_gs:
                if t_ is not None:
                    t_.record_stream(main)
        else:
            smooth_terms()
        stats = torch.empty(2 * S + 2, device=dev, dtype=torch.float64)
        _lib.call("fsnet_loss_finalize", acc, S, stats)
        ctx.S, ctx.cfg, ctx.flags, ctx.mdt = S, cfg, flags, mdt
        ctx.shapes = (B, H, W)
        ctx.unit_grads = (unit_gd, unit_gP)
        ctx.unit_gs = unit_gs
        ctx.colour = colour
        ctx.save_for_backward(*depths, *disps, tgt, packed, cam, P2c, acc, sums,
                              *( [ident] if ident is not None else []), *( [mask_c] if mask_c is not None else []),
                              *( [motion_c] if motion_c is not None else []), *[n for n in noise_c if n is not None])
        ctx.has = (ident is not None, mask_c is not None, motion_c is not None, noise_c[0] is not None)
        # the reference's loss is fp64 exactly when patched_mask is fp64
        total = stats[2 * S] if mdt == MASK_F64 else stats[2 * S].float()
        if sel is None:
            sel = torch.empty(0, device=dev, dtype=torch.uint8)
            pred0 = torch.empty(0, device=dev, dtype=torch.float32)
        ctx.mark_non_differentiable(stats, sel, pred0)
        return total.clone(), stats, sel, pred0

    @staticmethod
    def backward(ctx, g_total, _g_stats=None, _g_sel=None, _g_pred=None):
        S, cfg = ctx.S, ctx.cfg
        B, H, W = ctx.shapes
        sv = list(ctx.saved_tensors)
        depths, disps = sv[:S], sv[S:2 * S]
        tgt, packed, cam, P2c, acc, sums = sv[2 * S:2 * S + 6]
        rest = sv[2 * S + 6:]
        has_ident, has_mask, has_motion, has_noise = ctx.has
        ident = rest.pop(0) if has_ident else None
        mask_c = rest.pop(0) if has_mask else None
        motion_c = rest.pop(0) if has_motion else None
        noise_c = rest if has_noise else [None] * S
        dev = tgt.device
        gout = (g_total.detach().to(torch.float32) / S).reshape(1).contiguous()
        g_depths, g_disps = [], []
        gP = torch.zeros(B, 2, 12, device=dev, dtype=torch.float32) if ctx.need_pose else None
        mei = cfg.get("mei")
        warp_bwd, lut_args = ("fsnet_warp_ssim_bwd", ()) if mei is None else ("fsnet_warp_ssim_mei_bwd", (mei["lut"], mei["lut_idx"]))
        unit_gd, unit_gP = ctx.unit_grads
        gscale = g_total.detach().to(torch.float32)              # fused launches computed their gradients for d total / d loss = 1
        if gP is not None and unit_gP is not None:
            gP = unit_gP * gscale                                # the non-fused scales below accumulate into it
        for i, s in enumerate(cfg["scales"]):
            hs, ws = depths[i].shape[-2:]
            if ctx.fused[i]:
                g_depths.append(unit_gd[i] * gscale)
            else:
                full = (hs == H and ws == W)
                gd = (torch.empty if full else torch.zeros)(depths[i].shape, device=dev, dtype=torch.float32)
                _lib.call(warp_bwd, *lut_args, depths[i], hs, ws, packed, mask_c, ctx.mdt, cam, ident, noise_c[i],
                          motion_c, _lib.ctypes.c_uint(ctx.flags), B, H, W, acc[i], gout, gd, gP)
                g_depths.append(gd)
            if ctx.unit_gs[i] is not None:                       # computed under the forward's pair kernels for d total / d loss = 1
                g_disps.append(ctx.unit_gs[i] * gscale)
                continue
            h, w = disps[i].shape[-2:]
            gs = torch.empty(disps[i].shape, device=dev, dtype=torch.float32)
            _lib.call("fsnet_smooth_bwd", disps[i], ctx.colour[i], B, h, w, h, w, float(cfg["smooth_weight"] / (2 ** s)), sums[i], gout, gs)
            g_disps.append(gs)
        gT = [None, None]
        if gP is not None and mei is not None:
            # the kernel's 3x4 block IS cam_T_cam[:, :3, :4] (monodepth2_decoder.py:379-381); the last row gets no gradient
            for f in range(2):
                if ctx.needs_input_grad[2 + 2 * S + f]:
                    gT[f] = torch.nn.functional.pad(gP[:, f].reshape(B, 3, 4), (0, 0, 0, 1))
        elif gP is not None:
            # P = (K4 @ T)[:3]  =>  dL/dT = K4[:3,:]^T @ dL/dP      (Project3D, monodepth_utils.py:155)
            K3 = torch.zeros(B, 3, 4, device=dev, dtype=torch.float32)
            K3[:, :, :3] = P2c[:, :3, :3]
            for f in range(2):
                if ctx.needs_input_grad[2 + 2 * S + f]:
                    gT[f] = torch.matmul(K3.transpose(1, 2), gP[:, f].reshape(B, 3, 4))
        n_noise = S if has_noise else 0
        return (None, None, *g_depths, *g_disps, gT[0], gT[1], None, None, None, None, None, None, *([None] * n_noise))


def reprojection_loss(depths: Sequence[torch.Tensor], disps: Sequence[torch.Tensor], T0, T1, P2, tgt, src0, src1,
                      mask=None, motion=None, noise: Optional[Sequence[torch.Tensor]] = None, *, scales, overlapped_mask: bool,
                      smooth_weight: float = 1e-5, log_image: bool = False, mei: Optional[dict] = None, pre=None):
    """Returns (total, stats, sel, pred0) -- see _ReprojectionLoss; sel / pred0 are empty unless log_image.  ``noise[i]`` are standard-normal draws
    of shape [B,2,H,W] for scale i (monodepth2_decoder.py:258); None => no tie-break noise."""
    S = len(scales)
    cfg = dict(scales=list(scales), overlapped_mask=bool(overlapped_mask), smooth_weight=float(smooth_weight), log_image=log_image,
               mei=mei, pre=pre)
    extra = [] if noise is None else list(noise)
    return _ReprojectionLoss.apply(S, cfg, *depths, *disps, T0, T1, P2, tgt, src0, src1, mask, motion, *extra)


class _PoseMatrix(torch.autograd.Function):
    """transformation_from_parameters (monodepth_utils.py:46-63) as one launch per direction (csrc/smooth_head.cu)."""

    @staticmethod
    def forward(ctx, axisangle, translation, invert: bool):
        B = axisangle.shape[0]
        aa, tr = _f32c(axisangle).reshape(B, 3), _f32c(translation).reshape(B, 3)
        T = torch.empty(B, 4, 4, device=aa.device, dtype=torch.float32)
        _lib.call("fsnet_pose_matrix", aa, tr, B, int(bool(invert)), T)
        ctx.save_for_backward(aa, tr)
        ctx.invert, ctx.shapes = bool(invert), (axisangle.shape, translation.shape)
        return T

    @staticmethod
    def backward(ctx, gT):
        aa, tr = ctx.saved_tensors
        B = aa.shape[0]
        g_aa, g_tr = torch.empty_like(aa), torch.empty_like(tr)
        _lib.call("fsnet_pose_matrix_bwd", aa, tr, _f32c(gT), B, int(ctx.invert), g_aa, g_tr)
        return g_aa.view(ctx.shapes[0]), g_tr.view(ctx.shapes[1]), None


def pose_matrix(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False) -> torch.Tensor:
    """[B,1,3] axis-angle and translation -> [B,4,4] cam_T_cam."""
    return _PoseMatrix.apply(axisangle, translation, invert)


class MeiRayTable:
    """Device-resident cache of MeiCameraProjection.image2cam's look-up tables (mei_fisheye_utils.py:139-170).

    One ``[B,H,W,4]`` (X, Y, Z, mask) buffer whose slots remember the calibration they were built for;
    ``update`` launches ``fsnet_mei_lut`` (a plan kernel + an early-exit build kernel), so a calibration
    change is picked up on the device without the reference's per-sample ``.item()`` synchronisation and
    the call can sit inside a captured CUDA graph."""

    def __init__(self):
        self.key = None
        self.state = None
        self._calib_cache = {}

    @staticmethod
    def calib_host_tensor(calib_meta) -> torch.Tensor:
        """[B,3] fp64 (xi, k1, k2) on the host (pinned when CUDA is there): the graph-replayed hook copies it every step."""
        t = torch.tensor([(float(c["mirror_parameters"]["xi"]), float(c["distortion_parameters"]["k1"]),
                           float(c["distortion_parameters"]["k2"])) for c in calib_meta], dtype=torch.float64)
        return t.pin_memory() if torch.cuda.is_available() else t

    def calib_tensor(self, calib_meta, device):
        """[B,3] fp64 (xi, k1, k2) of the dataset's ``calib_meta`` dicts (fisheye_dataset.py:45-58,254)."""
        vals = tuple((float(c["mirror_parameters"]["xi"]), float(c["distortion_parameters"]["k1"]),
                      float(c["distortion_parameters"]["k2"])) for c in calib_meta)
        key = (vals, str(device))
        t = self._calib_cache.get(key)
        if t is None:
            if len(self._calib_cache) > 64:
                self._calib_cache.clear()
            t = torch.tensor(vals, dtype=torch.float64).to(device)
            self._calib_cache[key] = t
        return t

    def update(self, P2: torch.Tensor, calib: torch.Tensor, H: int, W: int) -> dict:
        B, dev = P2.shape[0], P2.device
        key = (B, H, W, str(dev))
        if self.key != key:
            self.key = key
            self.state = dict(header=torch.zeros(B, 8, device=dev, dtype=torch.float64),
                              lut_idx=torch.zeros(B, device=dev, dtype=torch.int32),
                              lut=torch.empty(B, H, W, 4, device=dev, dtype=torch.float32))
        st = self.state
        _lib.call("fsnet_mei_lut", _f32c(P2), calib, B, H, W, st["header"], st["lut_idx"], st["lut"])
        return dict(calib=calib, lut=st["lut"], lut_idx=st["lut_idx"])


def mei_depth(norm: torch.Tensor, mei: dict) -> torch.Tensor:
    """FishEyeDecoder.get_prediction (monodepth2_decoder.py:415-420): z of the back-projected ray."""
    B, _, H, W = norm.shape
    out = torch.empty_like(norm, dtype=torch.float32)
    _lib.call("fsnet_mei_depth", _f32c(norm), mei["lut"], mei["lut_idx"], B, H, W, out)
    return out


class _DepthHead(torch.autograd.Function):
    """softmax-over-bins -> depth -> disparity (depth_encoder.py:76-88,115-121) or the sigmoid head
    (:104-109).  logits may be NCHW-contiguous or channels_last."""

    @staticmethod
    def forward(ctx, logits, bins, scale, sigmoid_head: bool, min_depth: float, max_depth: float):
        B, n, h, w = logits.shape
        cl = logits.is_contiguous(memory_format=torch.channels_last) and not logits.is_contiguous()
        lg = logits.detach()
        if not cl:
            lg = lg.contiguous()
        dev = logits.device
        depth = torch.empty(B, 1, h, w, device=dev, dtype=torch.float32)
        disp = torch.empty(B, 1, h, w, device=dev, dtype=torch.float32)
        sc = None if scale is None else _f32c(scale).reshape(B)
        lib_args = (_RawPtr(lg), None if bins is None else _f32c(bins), sc, B, n, h, w, int(cl), int(sigmoid_head),
                    float(min_depth), float(max_depth))
        _lib.call("fsnet_depth_head_fwd", *lib_args, depth, disp)
        ctx.save_for_backward(lg, bins, sc)
        ctx.meta = (B, n, h, w, cl, sigmoid_head, min_depth, max_depth)
        return depth, disp

    @staticmethod
    def backward(ctx, g_depth, g_disp):
        lg, bins, sc = ctx.saved_tensors
        B, n, h, w, cl, sig, mn, mx = ctx.meta
        gl = torch.empty_like(lg)      # preserves the memory format
        gd = None if g_depth is None else _f32c(g_depth)
        gs = None if g_disp is None else _f32c(g_disp)
        _lib.call("fsnet_depth_head_bwd", _RawPtr(lg), None if bins is None else _f32c(bins), sc, B, n, h, w, int(cl), int(sig),
                  float(mn), float(mx), gd, gs, _RawPtr(gl))
        return gl, None, None, None, None, None


class _RawPtr:
    """Wraps a tensor whose memory is dense but not torch-'contiguous' (channels_last)."""

    def __init__(self, t):
        self.t = t


def depth_head(logits, bins, scale, sigmoid_head, min_depth, max_depth):
    return _DepthHead.apply(logits, bins, scale, sigmoid_head, min_depth, max_depth)


class _DistillLoss(torch.autograd.Function):
    """MonoDepth2Decoder.compute_distill_loss, scaled branch (monodepth2_decoder.py:185-203): one launch produces the mean
    and the unit gradients with respect to the student's depth and the un-activated uncertainty."""

    @staticmethod
    def forward(ctx, pred, teacher, ulogit):
        import ctypes
        p, t = _f32c(pred), _f32c(teacher)
        if p.shape != t.shape:
            raise _lib.FsnetError(f"distillation: student depth {tuple(p.shape)} vs teacher depth {tuple(t.shape)}")
        l = None if ulogit is None else _f32c(ulogit)
        if l is not None and l.numel() != p.numel():
            raise _lib.FsnetError(f"distillation: uncertainty map {tuple(l.shape)} vs depth map {tuple(p.shape)}")
        out = torch.zeros(1, device=p.device, dtype=torch.float64)
        gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        gu = torch.empty_like(l) if (l is not None and ctx.needs_input_grad[2]) else None
        _lib.call("fsnet_distill_loss", p, t, l, ctypes.c_longlong(p.numel()), out, gp, gu, None)
        ctx.gp, ctx.gu = gp, gu
        return out[0].float()

    @staticmethod
    def backward(ctx, g):
        g = g.float()
        gp = None if ctx.gp is None else (ctx.gp * g).view_as(ctx.gp)
        gu = None if ctx.gu is None else (ctx.gu * g).view_as(ctx.gu)
        return gp, None, gu


def distill_loss(pred: torch.Tensor, teacher: torch.Tensor, uncertain_logit: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean(|teacher - pred| / u + log(u + 1e-5)) with u = sigmoid(uncertain_logit); mean|teacher - pred| without it."""
    return _DistillLoss.apply(pred, teacher.detach(), uncertain_logit)
