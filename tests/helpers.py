"""Shared test helpers: build the B200 model through the reference-compatible plugin path."""
import numpy as np
from easydict import EasyDict as edict

from oracle import fsnet_oracle as O


def meta_arch_cfg(topo: O.Topology, is_log_image=False):
    head = edict(
        name="monodepth.networks.models.heads.monodepth2_decoder." + ("FishEyeDecoder" if topo.fisheye else "MonoDepth2Decoder"),
        scales=list(topo.scales), height=topo.height, width=topo.width, min_depth=topo.min_depth, max_depth=topo.max_depth,
        overlapped_mask=topo.overlapped_mask, is_log_image=is_log_image,
        depth_decoder_cfg=edict(
            name="monodepth.networks.models.heads.depth_encoder." + ("MultiChannelDepthDecoder" if topo.multi_channel else "DepthDecoder"),
            num_ch_enc=np.array(topo.num_ch_enc), num_output_channels=topo.n_bins, use_skips=topo.use_skips,
            scales=list(topo.scales), min_depth=topo.min_depth, max_depth=topo.max_depth, base_fx=topo.base_fx))
    backbone = edict(name="vision_base.networks.models.backbone.resnet.resnet", depth=topo.depth, pretrained=False,
                     frozen_stages=topo.frozen_stages, num_stages=4, out_indices=(-1, 0, 1, 2, 3), norm_eval=topo.norm_eval,
                     dilations=(1, 1, 1, 1))
    cfg = edict(depth_backbone_cfg=backbone, head_cfg=head, train_cfg=edict(frame_ids=list(topo.frame_ids)), test_cfg=edict())
    if topo.distill:
        import tempfile
        import torch
        t = O.teacher_topology(topo)
        head.distillation_loss_weight = topo.distill_weight
        head.is_uncertain_distill = topo.uncertain_distill
        head.depth_decoder_cfg.name = "monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoderUncertain"
        cfg.name = "monodepth.networks.models.meta_archs.monodepth2_model.DistillWPoseMeta"
        cfg.teacher_net_cfg = edict(
            name="monodepth.networks.models.meta_archs.teacher_model.MonoDepthInference", backbone_cfg=edict(backbone, depth=t.depth),
            depth_head_cfg=edict(name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
                                 num_ch_enc=np.array(t.num_ch_enc), num_output_channels=t.n_bins, use_skips=True, scales=list(t.scales),
                                 min_depth=t.min_depth, max_depth=t.max_depth))
        # the constructor reads a teacher checkpoint from disk (monodepth2_model.py:160-163)
        sd = O.make_state_dict(topo)
        with tempfile.NamedTemporaryFile(suffix=".pth", delete=False) as f:
            torch.save({k[len("teacher_net."):]: v for k, v in sd.items() if k.startswith("teacher_net.")}, f.name)
        cfg.teacher_net_path = f.name
    elif topo.posenet:
        cfg.name = "monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthMeta"
        cfg.pose_backbone_cfg = edict(backbone, depth=topo.pose_depth, num_input_images=2)
        head.pose_decoder_cfg = edict(name="monodepth.networks.models.heads.pose_decoder.PoseDecoder",
                                      num_ch_enc=np.array([64, 64, 128, 256, 512]), num_input_features=1, num_frames_to_predict_for=2)
    else:
        cfg.name = "monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthWPose"
    return cfg


def build_model(topo: O.Topology, seed=123, **kw):
    from vision_base.utils.builder import build
    cfg = meta_arch_cfg(topo, **kw)
    model = build(**cfg)
    if topo.distill:
        import os
        os.unlink(cfg.teacher_net_path)
    model.load_state_dict(O.make_state_dict(topo, seed), strict=True)
    return model.train()
