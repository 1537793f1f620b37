"""MonoDepth2Decoder: depth / pose heads + the photometric-reprojection loss.

Mirror of monodepth/networks/models/heads/monodepth2_decoder.py:19-347.  Constructor, attribute
conventions (every extra kwarg becomes an attribute read with ``getattr(self, name, default)``),
``forward_depth`` / ``forward_pose`` / ``loss`` / ``get_prediction`` and the returned
``{'loss','loss_dict','hm'}`` are the reference's.  The loss itself is NOT a PyTorch graph: it is one
fused CUDA kernel per scale plus small helpers (fsnet_b200/functional.py -> csrc/warp_ssim.cu).
"""
import os

import torch
import torch.nn as nn

from ..utils.builder import build
from .. import functional as Fn


class MonoDepth2Decoder(nn.Module):
    def __init__(self, scales, height, width, frame_ids, depth_decoder_cfg, pose_decoder_cfg=None,
                 multiscale_head_cfg=None, **kwargs):
        super().__init__()
        self.scales = scales
        self.num_scales = len(scales)
        self.height, self.width = height, width
        self.frame_ids = frame_ids
        self.depth_decoder = build(**depth_decoder_cfg)
        if pose_decoder_cfg is not None:
            self.pose_decoder = build(**pose_decoder_cfg)
        if multiscale_head_cfg is not None:
            self.multiscale_head = build(**multiscale_head_cfg)
        self.depth_metric_names = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]
        for key, value in kwargs.items():
            setattr(self, key, value)
        self.tie_break_noise = None     # parity tests inject the reference's randn draws here ({scale: [B,2,H,W]})

    def forward_pose(self, *args, **kwargs):
        return self.pose_decoder(*args, **kwargs)

    def forward_depth(self, features, *args, **kwargs):
        return self.depth_decoder(features, *args, **kwargs)

    def get_prediction(self, input_dict, output_dict):
        return dict(depth=output_dict[("depth", 0, 0)])

    def compute_pose_loss(self, output_dict, input_dict):
        """monodepth2_decoder.py:176-183 (weight 0 in every shipped config)."""
        loss = 0
        for f in self.frame_ids[1:]:
            loss = loss + torch.abs(input_dict[("relative_pose", f)] - output_dict[("cam_T_cam", f)]).mean()
        return loss

    def compute_distill_loss(self, output_dict, input_dict, scale):
        """monodepth2_decoder.py:185-203, scaled branch: |teacher - student| (divided by the predicted uncertainty, plus its
        log, when ``is_uncertain_distill``) averaged over the map -- one fused launch (csrc/distill.cu)."""
        ulogit = output_dict[("uncertain_logit", scale)] if getattr(self, "is_uncertain_distill", False) else None
        return Fn.distill_loss(output_dict[("depth", scale, scale)], output_dict[("teacher_depth", scale, scale)], ulogit)

    def _mei_camera(self, input_dict, H, W):
        """None for the pinhole camera; FishEyeDecoder returns the MEI ray table + calibration."""
        return None

    def _unsupported(self):
        for flag in ("is_residual_flow", "is_light_compensate", "learnable_photometric_uncertain", "is_ssim_weight"):
            if getattr(self, flag, False):
                raise NotImplementedError(f"{flag}=True is outside the B200 hot path (no shipped config enables it; "
                                          "the reference's own branch for it is incomplete, SURVEY.md 8(a) a7/a14)")
        if getattr(self, "residualflow_weight", 0) > 0:
            raise NotImplementedError("the residual-flow loss is outside the B200 hot path (no shipped config enables it)")
        if getattr(self, "distillation_loss_weight", 0) > 0 and getattr(self, "is_unscaled_distill", False):
            raise NotImplementedError("is_unscaled_distill=True is not implemented (no shipped config enables it)")

    def prefetch_loss_terms(self, input_dict):
        """Called by the meta-architectures at the START of a training step: everything the loss needs from the batch alone
        (identity photometric terms, packed frames, mask sums, colour pyramid, tie-break noise) is launched on a side stream and
        runs under the encoder (functional.prefetch_batch_terms).  Optional: without it the loss computes these itself."""
        self._pre = None
        tgt = input_dict.get(("original_image", 0))
        if (not self.training or os.environ.get("FSNET_LOSS_PREFETCH", "1") == "0" or len(self.frame_ids) != 3
                or tgt is None or tgt.device.type != "cuda" or not torch.is_grad_enabled()):
            return
        f1, f2 = self.frame_ids[1], self.frame_ids[2]
        draw = input_dict.get("motion_mask") is None and self.tie_break_noise is None
        self._pre = Fn.prefetch_batch_terms(tgt, input_dict[("original_image", f1)], input_dict[("original_image", f2)],
                                            input_dict.get("patched_mask"), self.scales, getattr(self, "is_log_image", True), draw)

    def compute_total_reprojection_loss(self, output_dict, input_dict):
        """Returns (losses, hm, total) like monodepth2_decoder.py:205-304."""
        if len(self.frame_ids) != 3:
            raise NotImplementedError("the fused loss kernel handles exactly two source frames (frame_ids=[0,a,b])")
        f1, f2 = self.frame_ids[1], self.frame_ids[2]
        tgt = input_dict[("original_image", 0)]
        B, _, H, W = tgt.shape
        depths = [output_dict[("depth", s, s)] for s in self.scales]
        disps = [output_dict[("disp", s)] for s in self.scales]
        motion = input_dict.get("motion_mask")
        pre, self._pre = getattr(self, "_pre", None), None
        noise = None
        if motion is None:
            if self.tie_break_noise is not None:
                noise = [self.tie_break_noise[s].to(tgt.device) for s in self.scales]
            elif pre is not None and pre.noise is not None:
                torch.cuda.current_stream().wait_event(pre.event)      # drawn on the side stream at step start
                noise = pre.noise
                for n_ in noise:
                    n_.record_stream(torch.cuda.current_stream())
            else:   # the reference draws on the CPU and uploads (:258-259); same distribution, drawn on the device
                noise = list(torch.randn(self.num_scales, B, 2, H, W, device=tgt.device).unbind(0))
        log_image = getattr(self, "is_log_image", True)
        mei = self._mei_camera(input_dict, H, W)
        total, stats, sel, pred0 = Fn.reprojection_loss(
            depths, disps, output_dict[("cam_T_cam", f1)], output_dict[("cam_T_cam", f2)], input_dict["P2"], tgt,
            input_dict[("original_image", f1)], input_dict[("original_image", f2)], input_dict.get("patched_mask"), motion,
            noise, scales=self.scales, overlapped_mask=getattr(self, "overlapped_mask", False), log_image=log_image, mei=mei, pre=pre)
        S = self.num_scales
        losses = {}
        for i, s in enumerate(self.scales):
            losses[f"smooth_loss/{s}"] = stats[S + i]
            losses[f"loss/{s}"] = stats[i]
        hm = {}
        if log_image and 0 in self.scales:
            hm["original_image"] = tgt[0:1]
            hm[f"predicted_image_{f1}"] = pred0[0:1]
            hm[f"predicted_image_{f2}"] = pred0[1:2]
            if motion is None:
                hm["loss_mask_0"] = dict(data=(sel[0:1] >= 2).unsqueeze(1))
        return losses, hm, total

    def loss(self, output_dict, input_dict):
        self._unsupported()
        losses, hm, total = self.compute_total_reprojection_loss(output_dict, input_dict)
        pose_weight = getattr(self, "pose_loss_weight", 0)
        if pose_weight > 0:
            pose_loss = self.compute_pose_loss(output_dict, input_dict)
            losses["pose_loss"] = pose_loss
            total = total + pose_weight * pose_loss
        distill_weight = getattr(self, "distillation_loss_weight", 0)
        if distill_weight > 0:                       # second training stage (monodepth2_decoder.py:328-334)
            for scale in self.scales:
                distill = self.compute_distill_loss(output_dict, input_dict, scale)
                losses[f"distilation/{scale}"] = distill.detach()
                total = total + distill * distill_weight
        losses["total_loss"] = total.detach()
        if not getattr(self, "is_log_image", True):
            hm = {}
        return {"loss": total, "loss_dict": losses, "hm": hm}


class FishEyeDecoder(MonoDepth2Decoder):
    """MonoDepth2Decoder for the KITTI-360 fisheye cameras (monodepth2_decoder.py:350-420): the decoder
    output is the NORM of the camera ray, back-projection goes through the MEI model's cached ray table
    (``inputs['calib_meta']`` + ``inputs['P2']``) and projection through ``cam2image`` with radial
    distortion; the overlap mask is additionally multiplied by the table's validity mask.  Same fused
    kernels as the pinhole head, instantiated for the MEI camera (csrc/warp_ssim.cu)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.mei_projection = Fn.MeiRayTable()

    def _mei_camera(self, input_dict, H, W):
        P2 = input_dict["P2"]
        calib = input_dict.get("calib_mei")          # optional pre-packed [B,3] fp64 (xi, k1, k2) device tensor
        if calib is None:
            calib = self.mei_projection.calib_tensor(input_dict["calib_meta"], P2.device)
        return self.mei_projection.update(P2, calib.double().contiguous(), H, W)

    def get_prediction(self, input_dict, output_dict):
        norm = output_dict[("depth", 0, 0)]
        mei = self._mei_camera(input_dict, norm.shape[-2], norm.shape[-1])
        return dict(depth=Fn.mei_depth(norm, mei), norm=norm)
