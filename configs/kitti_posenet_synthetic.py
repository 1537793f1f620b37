"""cfg2b "depth + pose": MonoDepthMeta with a 6-channel ResNet-18 PoseNet and PoseDecoder, wired as in the
reference's tests/example_cfgs/config.py:130-186; synthetic triplets, otherwise the KITTI recipe."""
import os
import sys

import numpy as np
from easydict import EasyDict as edict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kitti_wpose_synthetic import cfg as _base  # noqa: E402

sys.path.pop(0)
cfg = edict(_base)
cfg.path.project_path = cfg.path.project_path.replace("WPose", "PoseNet")
backbone = cfg.meta_arch.depth_backbone_cfg
head = edict(cfg.meta_arch.head_cfg)
head.pose_decoder_cfg = edict(name="monodepth.networks.models.heads.pose_decoder.PoseDecoder",
                              num_ch_enc=np.array([64, 64, 128, 256, 512]), num_input_features=1, num_frames_to_predict_for=2)
cfg.meta_arch = edict(
    name="monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthMeta",
    depth_backbone_cfg=backbone,
    pose_backbone_cfg=edict(backbone, num_input_images=2),
    head_cfg=head,
    train_cfg=edict(frame_ids=[0, 1, -1]),
    test_cfg=edict(),
)
