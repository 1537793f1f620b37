"""Meta-architectures (reference: vision_base/networks/models/meta_archs/base_meta.py:3-23 and
monodepth/networks/models/meta_archs/monodepth2_model.py:8-148)."""
from typing import Optional

import torch
import torch.nn as nn
from easydict import EasyDict

from ..utils.builder import build
from .pose_decoder import transformation_from_parameters


class BaseMetaArch(nn.Module):
    def forward_train(self, data, meta):
        raise NotImplementedError

    def forward_test(self, data, meta):
        raise NotImplementedError

    def dummy_forward(self, data):
        return dict()

    def forward(self, data, meta):
        return self.forward_train(data, meta) if meta["is_training"] else self.forward_test(data, meta)


class MonoDepthMeta(BaseMetaArch):
    """Depth net + PoseNet-predicted relative poses (monodepth2_model.py:8-64)."""

    def __init__(self, depth_backbone_cfg: EasyDict, pose_backbone_cfg: EasyDict, head_cfg: EasyDict,
                 train_cfg: EasyDict, test_cfg: EasyDict, **kwargs):
        super().__init__()
        self.depth_backbone = build(**depth_backbone_cfg)
        self.pose_backbone = build(**pose_backbone_cfg)
        self.head = build(frame_ids=train_cfg.frame_ids, **head_cfg)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    def forward_train(self, data, meta):
        self.head.prefetch_loss_terms(data)
        features = self.depth_backbone(data[("image", 0)])
        outputs = self.head.forward_depth(features)
        for f_i in self.train_cfg.frame_ids[1:]:
            pair = [data[("image", f_i)], data[("image", 0)]] if f_i < 0 else [data[("image", 0)], data[("image", f_i)]]
            axisangle, translation = self.head.forward_pose([self.pose_backbone(torch.cat(pair, 1))])
            outputs[("axisangle", f_i)] = axisangle
            outputs[("translation", f_i)] = translation
            outputs[("cam_T_cam", f_i)] = transformation_from_parameters(axisangle[:, 0], translation[:, 0], invert=(f_i < 0))
        return self.head.loss(outputs, data)

    def dummy_forward(self, image):
        outputs = self.head.forward_depth(self.depth_backbone(image))
        return self.head.get_prediction(None, outputs)

    def forward_test(self, data, meta):
        outputs = self.head.forward_depth(self.depth_backbone(data[("image", 0)]))
        return self.head.get_prediction(data, outputs)


class MonoDepthWPose(BaseMetaArch):
    """Depth net with dataset ("given") poses (monodepth2_model.py:66-148).  The reference's optional
    residual-pose branch needs a ``PoseDecoder.forward(features, base_pose)`` that its repository does
    not contain; asking for it raises instead of failing later with a TypeError."""

    def __init__(self, depth_backbone_cfg: EasyDict, head_cfg: EasyDict, train_cfg: EasyDict, test_cfg: EasyDict,
                 pose_backbone_cfg: Optional[EasyDict] = None, **kwargs):
        super().__init__()
        self.depth_backbone = build(**depth_backbone_cfg)
        self.head = build(frame_ids=train_cfg.frame_ids, **head_cfg)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.is_use_res_pose = pose_backbone_cfg is not None
        if self.is_use_res_pose:
            raise NotImplementedError("MonoDepthWPose with pose_backbone_cfg: the residual-pose decoder is not part of the "
                                      "reference repository (its shipped PoseDecoder rejects the call)")

    def forward_train(self, data, meta):
        self.head.prefetch_loss_terms(data)
        features = self.depth_backbone(data[("image", 0)])
        outputs = self.head.forward_depth(features, data["P2"])
        for f_i in self.train_cfg.frame_ids[1:]:
            outputs[("cam_T_cam", f_i)] = data[("relative_pose", f_i)]
        return self.head.loss(outputs, data)

    def forward_test(self, data, meta):
        outputs = self.head.forward_depth(self.depth_backbone(data[("image", 0)]), data["P2"])
        return self.head.get_prediction(data, outputs)

    def dummy_forward(self, image):
        outputs = self.head.forward_depth(self.depth_backbone(image))
        return self.head.get_prediction(None, outputs)


class MonoDepthInference(nn.Module):
    """The frozen teacher of the distillation stage: encoder + depth decoder, called without P2
    (monodepth/networks/models/meta_archs/teacher_model.py:5-32)."""

    def __init__(self, backbone_cfg: EasyDict, depth_head_cfg: EasyDict, is_produce_detached: bool = True, **kwargs):
        super().__init__()
        self.depth_backbone = build(**backbone_cfg)
        self.depth_decoder = build(**depth_head_cfg)
        self.is_produce_detached = is_produce_detached

    def forward(self, x):
        return self.depth_decoder(self.depth_backbone(x))

    def compute_teacher_depth(self, x):
        if self.is_produce_detached:
            with torch.no_grad():
                out = self(x)
        else:
            out = self(x)
        return {("teacher_depth", k[1], k[2]): v for k, v in out.items() if k[0] == "depth"}


class DistillWPoseMeta(BaseMetaArch):
    """Second training stage (monodepth2_model.py:150-206): the stage-1 network, loaded from ``teacher_net_path`` and kept
    frozen in eval mode, supervises a student with uncertainty heads next to the photometric loss; dataset poses."""

    def __init__(self, teacher_net_cfg: EasyDict, depth_backbone_cfg: EasyDict, teacher_net_path: str, head_cfg: EasyDict,
                 train_cfg: EasyDict, test_cfg: EasyDict, **kwargs):
        super().__init__()
        self.teacher_net = build(**teacher_net_cfg)
        self.teacher_net.load_state_dict(torch.load(teacher_net_path, map_location="cpu"), strict=False)
        for p in self.teacher_net.parameters():
            p.requires_grad = False
        self.depth_backbone = build(**depth_backbone_cfg)
        self.head = build(frame_ids=train_cfg.frame_ids, **head_cfg)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    def train(self, mode=True):
        super().train(mode)
        self.teacher_net.eval()            # running statistics, never updated
        return self

    def forward_train(self, data, meta):
        self.head.prefetch_loss_terms(data)
        image_0 = data[("image", 0)]
        outputs = self.head.forward_depth(self.depth_backbone(image_0), data["P2"])
        outputs.update(self.teacher_net.compute_teacher_depth(image_0))
        for f_i in self.train_cfg.frame_ids[1:]:
            outputs[("cam_T_cam", f_i)] = data[("relative_pose", f_i)]
        return self.head.loss(outputs, data)

    def forward_test(self, data, meta):
        outputs = self.head.forward_depth(self.depth_backbone(data[("image", 0)]), data["P2"])
        return self.head.get_prediction(data, outputs)

    def dummy_forward(self, image):
        outputs = self.head.forward_depth(self.depth_backbone(image))
        return self.head.get_prediction(None, outputs)
