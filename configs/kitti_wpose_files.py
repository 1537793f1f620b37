"""The reference's configs/kitti_wpose_example recipe on KITTI raw FILES (readers + augmentation lists + model + hooks all
through the reference's dotted names); only the path entries differ: they come from the environment.
    FSNET_KITTI_PATH   KITTI raw root (dates / drives / image_02, image_03, oxts/pose.mat, calibration files)
    FSNET_KITTI_SPLIT  training split file (default <repo>/meta_data/eigen_zhou/train_files.txt)
    FSNET_KITTI_VAL_SPLIT  evaluation split file (default = training split)
    FSNET_DEVICE_AUG   1: augment on the GPU (uint8 frames + drawn parameters are uploaded; same list, same draws)
    FSNET_KITTI_GT     ground-truth export (npz) of the evaluation split: when set, the reference's evaluate_hook
                       (KittiEvaluationHook + KittiEigenEvaluator) is configured and runs every FSNET_TEST_ITER (5) epochs;
                       the file is written from the split's Velodyne scans on first use
"""
import os

import numpy as np
from easydict import EasyDict as edict

cfg = edict()
path = edict()
path.base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if "__file__" in globals() else os.getcwd()
path.kitti_path = os.environ.get("FSNET_KITTI_PATH", "/data/kitti_raw")
path.project_path = os.path.join(os.environ.get("FSNET_WORKDIR", "/tmp/fsnet_b200_workdirs"), "Kitti_MonoDepth2WPose")
path.log_path = os.path.join(path.project_path, "log")
path.checkpoint_path = os.path.join(path.project_path, "checkpoint")
for _p in (path.project_path, path.log_path, path.checkpoint_path):
    os.makedirs(_p, exist_ok=True)
cfg.path = path

cfg.trainer = edict(
    gpu=0, max_epochs=20, disp_iter=50, save_iter=5, test_iter=0,
    training_hook=edict(name="vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=35.0),
)
_val_split = os.environ.get("FSNET_KITTI_VAL_SPLIT", os.environ.get("FSNET_KITTI_SPLIT", os.path.join(path.base_path, "meta_data", "eigen", "test_files.txt")))
if os.environ.get("FSNET_KITTI_GT"):
    cfg.trainer.test_iter = int(os.environ.get("FSNET_TEST_ITER", 5))
    cfg.trainer.evaluate_hook = edict(
        name="monodepth.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.KittiEvaluationHook",
        test_run_hook_cfg=edict(name="vision_base.pipeline_hooks.train_val_hooks.base_validation_hooks.BaseValidationHook"),
        dataset_eval_cfg=edict(name="monodepth.evaluation.kitti_unsupervised_eval.KittiEigenEvaluator", data_path=path.kitti_path,
                               split_file=_val_split, gt_saved_file=os.environ["FSNET_KITTI_GT"]),
    )
cfg.optimizer = edict(name="adam", lr=1e-4, weight_decay=0)
cfg.scheduler = edict(name="StepLR", step_size=15)

data = edict(batch_size=12, num_workers=4, rgb_shape=(192, 640, 3), frame_idxs=[0, 1, -1])
split = os.environ.get("FSNET_KITTI_SPLIT", os.path.join(path.base_path, "meta_data", "eigen_zhou", "train_files.txt"))
train_dataset = edict(
    name="vision_base.data.datasets.dataset_utils.ConcatDataset", frame_idxs=data.frame_idxs, is_motion_mask=False,
    is_precompute_flow=False, is_filter_static=True,
    cfg_list=[edict(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoDataset", raw_path=path.kitti_path, split_file=split)],
)
val_dataset = edict(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoEigenTestDataset", raw_path=path.kitti_path,
                    split_file=os.environ.get("FSNET_KITTI_VAL_SPLIT", split))

resize_keys = [("image", i) for i in data.frame_idxs] + [("original_image", i) for i in data.frame_idxs]
color_keys = [("image", i) for i in data.frame_idxs]
original_keys = [("original_image", i) for i in data.frame_idxs]
pose_axis_pairs = [(("relative_pose", i), 0) for i in data.frame_idxs[1:]]
data.augmentation = edict(rgb_mean=np.array([0.485, 0.456, 0.406]), rgb_std=np.array([0.229, 0.224, 0.225]),
                          cropSize=(data.rgb_shape[0], data.rgb_shape[1]),
                          key_mappings=edict(image_keys=resize_keys, calib_keys=["P2"], gt_image_keys=["patched_mask"]))
A = "vision_base.data.augmentations.augmentations"
train_dataset.augmentation = edict(
    name="vision_base.utils.builder.Sequential",
    cfg_list=[
        edict(name=f"{A}.ConvertToFloat"),
        edict(name=f"{A}.RandomWarpAffine", output_w=data.augmentation.cropSize[1], output_h=data.augmentation.cropSize[0],
              shift_border=int(os.environ.get("FSNET_SHIFT_BORDER", 128))),
        edict(name=f"{A}.RandomMirror", mirror_prob=0.5, pose_axis_pairs=pose_axis_pairs),
        edict(name="vision_base.utils.builder.Shuffle", image_keys=color_keys, cfg_list=[
            edict(name=f"{A}.RandomBrightness", distort_prob=1.0),
            edict(name=f"{A}.RandomContrast", distort_prob=1.0, lower=0.6, upper=1.4),
            edict(name="vision_base.utils.builder.Sequential", cfg_list=[
                edict(name=f"{A}.ConvertColor", transform="HSV"),
                edict(name=f"{A}.RandomSaturation", distort_prob=1.0, lower=0.6, upper=1.4),
                edict(name=f"{A}.ConvertColor", current="HSV", transform="RGB"),
            ]),
        ]),
        edict(name=f"{A}.Normalize", mean=data.augmentation.rgb_mean, stds=data.augmentation.rgb_std, image_keys=color_keys),
        edict(name=f"{A}.Normalize", mean=np.array([0, 0, 0]), stds=np.array([1, 1, 1]), image_keys=original_keys),
        edict(name=f"{A}.ConvertToTensor"),
    ],
    **data.augmentation.key_mappings,
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
        name="monodepth.networks.models.heads.monodepth2_decoder.MonoDepth2Decoder",
        scales=[0, 1, 2, 3], height=data.rgb_shape[0], width=data.rgb_shape[1], min_depth=0.5, max_depth=100.0,
        overlapped_mask=True, is_log_image=False,
        depth_decoder_cfg=edict(
            name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
            num_ch_enc=np.array([64, 64, 128, 256, 512]), num_output_channels=16, use_skips=True, scales=[0, 1, 2, 3],
            min_depth=0.5, max_depth=100)),
    train_cfg=edict(frame_ids=[0, 1, -1]),
    test_cfg=edict(),
)
