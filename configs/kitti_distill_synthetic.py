"""Second training stage: the reference's configs/distill_kitti_example topology (DistillWPoseMeta: frozen stage-1 teacher,
ResNet-18 student with MultiChannelDepthDecoderUncertain, distillation weight 0.3 with predicted uncertainty) on the
synthetic triplet dataset.  The teacher checkpoint comes from stage 1:
    python scripts/train.py --config=configs/kitti_wpose_synthetic.py ...
    python monodepth/transform_teacher.py <..._latest.pth> teacher.pth
    FSNET_TEACHER=teacher.pth python scripts/train.py --config=configs/kitti_distill_synthetic.py ..."""
import os

import numpy as np
from easydict import EasyDict as edict

cfg = edict()

path = edict()
path.base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if "__file__" in globals() else os.getcwd()
path.project_path = os.path.join(os.environ.get("FSNET_WORKDIR", "/tmp/fsnet_b200_workdirs"), "Kitti_distill_synthetic")
path.log_path = os.path.join(path.project_path, "log")
path.checkpoint_path = os.path.join(path.project_path, "checkpoint")
for _p in (path.project_path, path.log_path, path.checkpoint_path):
    os.makedirs(_p, exist_ok=True)
cfg.path = path

cfg.trainer = edict(
    gpu=0, max_epochs=5, disp_iter=50, save_iter=5, test_iter=0,
    training_hook=edict(name="vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=35.0),
)
cfg.optimizer = edict(name="adam", lr=1e-4, weight_decay=0)
cfg.scheduler = edict(name="StepLR", step_size=4)

data = edict(batch_size=12, num_workers=2, rgb_shape=(192, 640, 3), frame_idxs=[0, 1, -1])
cfg.data = data
cfg.train_dataset = edict(name="vision_base.data.datasets.synthetic.SyntheticTripletDataset", length=1200,
                          height=data.rgb_shape[0], width=data.rgb_shape[1], frame_idxs=data.frame_idxs)
cfg.val_dataset = edict(name="vision_base.data.datasets.synthetic.SyntheticTripletDataset", length=16,
                        height=data.rgb_shape[0], width=data.rgb_shape[1], frame_idxs=[0])

_backbone = edict(name="vision_base.networks.models.backbone.resnet.resnet", depth=18, pretrained=False, frozen_stages=-1,
                  num_stages=4, out_indices=(-1, 0, 1, 2, 3), norm_eval=False, dilations=(1, 1, 1, 1))
_decoder = dict(num_ch_enc=np.array([64, 64, 128, 256, 512]), num_output_channels=16, use_skips=True, scales=[0, 1, 2, 3],
                min_depth=0.5, max_depth=100)
cfg.meta_arch = edict(
    name="monodepth.networks.models.meta_archs.monodepth2_model.DistillWPoseMeta",
    teacher_net_cfg=edict(
        name="monodepth.networks.models.meta_archs.teacher_model.MonoDepthInference", backbone_cfg=edict(_backbone),
        depth_head_cfg=edict(name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder", **_decoder)),
    teacher_net_path=os.environ.get("FSNET_TEACHER", os.path.join(path.base_path, "kitti_teacher.pth")),
    depth_backbone_cfg=edict(_backbone),
    head_cfg=edict(
        name="monodepth.networks.models.heads.monodepth2_decoder.MonoDepth2Decoder",
        scales=[0, 1, 2, 3], height=data.rgb_shape[0], width=data.rgb_shape[1], min_depth=0.5, max_depth=100.0,
        overlapped_mask=True, is_log_image=False, distillation_loss_weight=0.3, is_uncertain_distill=True,
        depth_decoder_cfg=edict(name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoderUncertain", **_decoder)),
    train_cfg=edict(frame_ids=[0, 1, -1]),
    test_cfg=edict(),
)
