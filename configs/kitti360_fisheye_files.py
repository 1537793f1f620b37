"""The reference's configs/kitti360_fisheye_example recipe on KITTI-360 FILES (fisheye reader + augmentation list + FishEyeDecoder
through the reference's dotted names); only the path entries differ: they come from the environment.
    FSNET_KITTI360_PATH   KITTI-360 root (calibration/, data_poses/, data_2d_raw/)
    FSNET_KITTI360_SPLIT  meta file (sequence,pose_index,image_index,former,latter per line)
    FSNET_FISHEYE_MASK    optional validity-mask image
(The shipped reference config references an undefined `color_augmented_image_keys`; the obvious
definition -- the ('image', f) keys -- is used here.)
"""
import os

import numpy as np
from easydict import EasyDict as edict

cfg = edict()
path = edict()
path.base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if "__file__" in globals() else os.getcwd()
path.kitti360_path = os.environ.get("FSNET_KITTI360_PATH", "/data/KITTI-360")
path.project_path = os.path.join(os.environ.get("FSNET_WORKDIR", "/tmp/fsnet_b200_workdirs"), "Kitti360_fisheye")
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

_size = int(os.environ.get("FSNET_FISHEYE_SIZE", 384))
data = edict(batch_size=16, num_workers=4, rgb_shape=(_size, _size, 3), frame_idxs=[0, 1, -1])
split = os.environ.get("FSNET_KITTI360_SPLIT", os.path.join(path.base_path, "meta_data", "kitti360_trainsub", "kitti360_train.txt"))
_mask = os.environ.get("FSNET_FISHEYE_MASK")
_ds = edict(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=path.kitti360_path, split_file=split)
if _mask:
    _ds.fisheye_mask = _mask
train_dataset = edict(name="vision_base.data.datasets.dataset_utils.ConcatDataset", frame_idxs=data.frame_idxs, is_motion_mask=False,
                      is_precompute_flow=False, is_filter_static=True, cfg_list=[_ds])
val_dataset = edict(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=path.kitti360_path,
                    split_file=os.environ.get("FSNET_KITTI360_VAL_SPLIT", split), is_filter_static=False, use_right_image=False)

image_keys = [("image", i) for i in data.frame_idxs]
original_keys = [("original_image", i) for i in data.frame_idxs]
pose_axis_pairs = [(("relative_pose", i), 0) for i in data.frame_idxs[1:]]
data.augmentation = edict(rgb_mean=np.array([0.485, 0.456, 0.406]), rgb_std=np.array([0.229, 0.224, 0.225]),
                          cropSize=(data.rgb_shape[0], data.rgb_shape[1]))
A = "vision_base.data.augmentations.augmentations"
train_dataset.augmentation = edict(
    name="vision_base.utils.builder.Sequential",
    cfg_list=[
        edict(name=f"{A}.ConvertToFloat"),
        edict(name=f"{A}.Resize", size=data.augmentation.cropSize, preserve_aspect_ratio=True, force_pad=True),
        edict(name=f"{A}.RandomMirror", mirror_prob=0.5, pose_axis_pairs=pose_axis_pairs),
        edict(name=f"{A}.Copy", from_keys=image_keys, to_keys=original_keys),
        edict(name="vision_base.utils.builder.Shuffle", cfg_list=[
            edict(name=f"{A}.RandomBrightness", distort_prob=1.0),
            edict(name=f"{A}.RandomContrast", distort_prob=1.0, lower=0.6, upper=1.4),
            edict(name="vision_base.utils.builder.Sequential", cfg_list=[
                edict(name=f"{A}.ConvertColor", transform="HSV"),
                edict(name=f"{A}.RandomSaturation", distort_prob=1.0, lower=0.6, upper=1.4),
                edict(name=f"{A}.ConvertColor", current="HSV", transform="RGB"),
            ]),
        ]),
        edict(name=f"{A}.Normalize", mean=data.augmentation.rgb_mean, stds=data.augmentation.rgb_std, image_keys=image_keys),
        edict(name=f"{A}.Normalize", mean=np.array([0, 0, 0]), stds=np.array([1, 1, 1]), image_keys=original_keys),
        edict(name=f"{A}.ConvertToTensor", image_keys=image_keys + original_keys),
    ],
    image_keys=image_keys, calib_keys=["P2"], gt_image_keys=["patched_mask"],
)
if int(os.environ.get("FSNET_DEVICE_AUG", "0")):
    # same list, same random draws; the loader ships uint8 frames and the GPU does the pixel work (fsnet_b200/data/device_augment.py)
    train_dataset.augmentation = edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=train_dataset.augmentation)
val_dataset.augmentation = edict(
    name="vision_base.utils.builder.Sequential",
    cfg_list=[
        edict(name=f"{A}.ConvertToFloat"),
        edict(name=f"{A}.Resize", size=data.augmentation.cropSize, preserve_aspect_ratio=False),
        edict(name=f"{A}.Normalize", mean=data.augmentation.rgb_mean, stds=data.augmentation.rgb_std),
        edict(name=f"{A}.ConvertToTensor"),
    ],
    image_keys=[("image", 0)], calib_keys=["P2"],
)
cfg.data = data
cfg.train_dataset = train_dataset
cfg.val_dataset = val_dataset

cfg.meta_arch = edict(
    name="monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthWPose",
    depth_backbone_cfg=edict(
        name="vision_base.networks.models.backbone.resnet.resnet", depth=18,
        pretrained=bool(int(os.environ.get("FSNET_PRETRAINED", "0"))), frozen_stages=-1,
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
