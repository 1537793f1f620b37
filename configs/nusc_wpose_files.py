"""The reference's configs/nusc_wpose_example recipe on nuScenes JSON exports (reader + augmentation list + model + hooks all
through the reference's dotted names); only the path entries differ: they come from the environment.
    FSNET_NUSC_JSON      comma-separated training JSON exports (the reference concatenates the key-frame and the sweep export)
    FSNET_NUSC_VAL_JSON  evaluation JSON export (default = the first training export)
    FSNET_NUSC_SIZE      HxW of the network input (default 288x512, the reference's)
The reference's evaluate_hook (FastNuscEvaluationHook + NuscenesEvaluator: LiDAR ground-truth generation through the nuScenes
devkit) is outside SURVEY.md section 8 and is not configured here.
"""
import os

import numpy as np
from easydict import EasyDict as edict

cfg = edict()
path = edict()
path.base_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if "__file__" in globals() else os.getcwd()
path.project_path = os.path.join(os.environ.get("FSNET_WORKDIR", "/tmp/fsnet_b200_workdirs"), "nusc_wpose")
path.log_path = os.path.join(path.project_path, "log")
path.checkpoint_path = os.path.join(path.project_path, "checkpoint")
for _p in (path.project_path, path.log_path, path.checkpoint_path):
    os.makedirs(_p, exist_ok=True)
cfg.path = path

cfg.trainer = edict(
    gpu=0, max_epochs=10, disp_iter=50, save_iter=5, test_iter=0,
    training_hook=edict(name="vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=1.0),
)
cfg.optimizer = edict(name="adam", lr=1e-4, weight_decay=0)
cfg.scheduler = edict(name="StepLR", step_size=4)

_h, _w = (int(v) for v in os.environ.get("FSNET_NUSC_SIZE", "288x512").lower().split("x"))
data = edict(batch_size=8, num_workers=4, rgb_shape=(_h, _w, 3), frame_idxs=[0, 1, -1])
_meta = os.path.join(path.base_path, "meta_data", "nusc_trainsub")
_jsons = os.environ.get("FSNET_NUSC_JSON", ",".join(os.path.join(_meta, f) for f in ("json_nusc_front_train.json", "json_nusc_sweep_train.json")))
_jsons = [j for j in _jsons.split(",") if j]
train_dataset = edict(
    name="vision_base.data.datasets.dataset_utils.ConcatDataset", frame_idxs=data.frame_idxs, is_motion_mask=False,
    is_precompute_flow=False, is_filter_static=True,
    cfg_list=[edict(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset", json_path=j) for j in _jsons],
)
val_dataset = edict(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset",
                    json_path=os.environ.get("FSNET_NUSC_VAL_JSON", _jsons[0]), image_keys=["frame0"], frame_ids=[0])

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
        edict(name=f"{A}.Resize", size=data.augmentation.cropSize, preserve_aspect_ratio=True, force_pad=True),
        edict(name="vision_base.utils.builder.Shuffle", image_keys=color_keys, cfg_list=[
            edict(name=f"{A}.RandomBrightness", distort_prob=1.0),
            edict(name=f"{A}.RandomContrast", distort_prob=1.0, lower=0.6, upper=1.4),
            edict(name="vision_base.utils.builder.Sequential", cfg_list=[
                edict(name=f"{A}.ConvertColor", transform="HSV"),
                edict(name=f"{A}.RandomSaturation", distort_prob=1.0, lower=0.6, upper=1.4),
                edict(name=f"{A}.ConvertColor", current="HSV", transform="RGB"),
            ]),
        ]),
        edict(name=f"{A}.RandomMirror", mirror_prob=0.5, pose_axis_pairs=pose_axis_pairs),
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
        name="vision_base.networks.models.backbone.resnet.resnet", depth=int(os.environ.get("FSNET_NUSC_DEPTH", 34)),
        pretrained=bool(int(os.environ.get("FSNET_PRETRAINED", "0"))), frozen_stages=-1,
        num_stages=4, out_indices=(-1, 0, 1, 2, 3), norm_eval=False, dilations=(1, 1, 1, 1)),
    head_cfg=edict(
        name="monodepth.networks.models.heads.monodepth2_decoder.MonoDepth2Decoder",
        scales=[0, 1, 2, 3], height=data.rgb_shape[0], width=data.rgb_shape[1], min_depth=0.5, max_depth=100.0,
        overlapped_mask=False, is_log_image=False,
        depth_decoder_cfg=edict(
            name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
            num_ch_enc=np.array([64, 64, 128, 256, 512]), num_output_channels=64, use_skips=True, scales=[0, 1, 2, 3],
            min_depth=0.5, max_depth=100, base_fx=369)),
    train_cfg=edict(frame_ids=[0, 1, -1]),
    test_cfg=edict(),
)
