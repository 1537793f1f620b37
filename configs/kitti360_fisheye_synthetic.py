"""cfg5: the reference's configs/kitti360_fisheye_example topology (FishEyeDecoder on the MEI camera model,
ResNet-18, 64 depth bins, max_depth 150, overlapped mask, is_log_image unset => True, clip 1.0, StepLR(8))
at BASELINE.json's 512x512 with only the dataset and path entries edited: the KITTI-360 reader is replaced
by the synthetic fisheye triplet dataset (KITTI-360-like MEI calibration scaled to the crop)."""
import os

import numpy as np
from easydict import EasyDict as edict

cfg = edict()

path = edict()
path.base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if "__file__" in globals() else os.getcwd()
path.project_path = os.path.join(os.environ.get("FSNET_WORKDIR", "/tmp/fsnet_b200_workdirs"), "Kitti360_fisheye_synthetic")
path.log_path = os.path.join(path.project_path, "log")
path.checkpoint_path = os.path.join(path.project_path, "checkpoint")
for _p in (path.project_path, path.log_path, path.checkpoint_path):
    os.makedirs(_p, exist_ok=True)
cfg.path = path

cfg.trainer = edict(
    gpu=0, max_epochs=20, disp_iter=50, save_iter=5, test_iter=0,
    training_hook=edict(name="vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=1.0),
)
cfg.optimizer = edict(name="adam", lr=1e-4, weight_decay=0)
cfg.scheduler = edict(name="StepLR", step_size=8)

data = edict(batch_size=4, num_workers=2, rgb_shape=(512, 512, 3), frame_idxs=[0, 1, -1])
cfg.data = data
cfg.train_dataset = edict(name="vision_base.data.datasets.synthetic.SyntheticTripletDataset", length=1200,
                          height=data.rgb_shape[0], width=data.rgb_shape[1], frame_idxs=data.frame_idxs, fisheye=True)
cfg.val_dataset = edict(name="vision_base.data.datasets.synthetic.SyntheticTripletDataset", length=16,
                        height=data.rgb_shape[0], width=data.rgb_shape[1], frame_idxs=[0], fisheye=True)

cfg.meta_arch = edict(
    name="monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthWPose",
    depth_backbone_cfg=edict(
        name="vision_base.networks.models.backbone.resnet.resnet", depth=18, pretrained=False, frozen_stages=-1,
        num_stages=4, out_indices=(-1, 0, 1, 2, 3), norm_eval=False, dilations=(1, 1, 1, 1)),
    head_cfg=edict(
        name="monodepth.networks.models.heads.monodepth2_decoder.FishEyeDecoder",
        scales=[0, 1, 2, 3], height=data.rgb_shape[0], width=data.rgb_shape[1], min_depth=0.5, max_depth=150.0,
        overlapped_mask=True,
        depth_decoder_cfg=edict(
            name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
            num_ch_enc=np.array([64, 64, 128, 256, 512]), num_output_channels=64, use_skips=True, scales=[0, 1, 2, 3],
            min_depth=0.5, max_depth=150)),
    train_cfg=edict(frame_ids=[0, 1, -1]),
    test_cfg=edict(),
)
