"""Shared by tests/golden/make_golden_aug.py (reference side) and tests/test_augmentations_cpu.py (this repo's side):
a raw dataset-style sample and the augmentation lists of configs/kitti_wpose_example:108-172, written with dotted names
that resolve in whichever `vision_base` is first on sys.path."""
import numpy as np
from easydict import EasyDict as edict

FRAMES = [0, 1, -1]
AUG = "vision_base.data.augmentations.augmentations"
OUT_H, OUT_W = 48, 160


def raw_sample(seed, h=120, w=400):
    """What mono_dataset.py:179-218 hands to the transform: uint8 RGB frames, their 'original_image' copies, P2,
    relative poses and an all-ones patched mask."""
    g = np.random.default_rng(seed)
    data = {}
    for f in FRAMES:
        lo = g.integers(0, 256, size=(h // 8, w // 8, 3)).astype(np.uint8)
        img = np.kron(lo, np.ones((8, 8, 1), dtype=np.uint8)) // 2 + g.integers(0, 128, size=(h, w, 3)).astype(np.uint8)
        data[("image", f)] = img
        data[("original_image", f)] = img.copy()
    P2 = np.zeros((3, 4), dtype=np.float32)
    P2[0, 0], P2[1, 1], P2[0, 2], P2[1, 2], P2[2, 2], P2[0, 3] = 0.58 * w, 1.92 * h, 0.5 * w, 0.5 * h, 1.0, 4.5
    data["P2"] = P2
    data["original_P2"] = P2.copy()
    for f in FRAMES[1:]:
        ang = g.uniform(-0.05, 0.05, size=3)
        cx, cy, cz = np.cos(ang)
        sx, sy, sz = np.sin(ang)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = (Rz @ Ry @ Rx).astype(np.float32)
        T[:3, 3] = g.uniform(-1, 1, size=3).astype(np.float32)
        data[("relative_pose", f)] = T
    data["patched_mask"] = np.ones([h, w])
    return data


def _keys():
    resize = [("image", i) for i in FRAMES] + [("original_image", i) for i in FRAMES]
    color = [("image", i) for i in FRAMES]
    return resize, color, edict(image_keys=resize, calib_keys=["P2"], gt_image_keys=["patched_mask"])


def train_cfg():
    resize, color, mappings = _keys()
    poses = [(("relative_pose", i), 0) for i in FRAMES[1:]]
    return edict(name="vision_base.utils.builder.Sequential", cfg_list=[
        edict(name=f"{AUG}.ConvertToFloat"),
        edict(name=f"{AUG}.RandomWarpAffine", output_w=OUT_W, output_h=OUT_H, shift_border=32),
        edict(name=f"{AUG}.RandomMirror", mirror_prob=0.5, pose_axis_pairs=poses),
        edict(name="vision_base.utils.builder.Shuffle", cfg_list=[
            edict(name=f"{AUG}.RandomBrightness", distort_prob=1.0),
            edict(name=f"{AUG}.RandomContrast", distort_prob=1.0, lower=0.6, upper=1.4),
            edict(name="vision_base.utils.builder.Sequential", cfg_list=[
                edict(name=f"{AUG}.ConvertColor", transform="HSV"),
                edict(name=f"{AUG}.RandomSaturation", distort_prob=1.0, lower=0.6, upper=1.4),
                edict(name=f"{AUG}.ConvertColor", current="HSV", transform="RGB"),
            ]),
        ], image_keys=color),
        edict(name=f"{AUG}.Normalize", mean=np.array([0.485, 0.456, 0.406]), stds=np.array([0.229, 0.224, 0.225]), image_keys=color),
        edict(name=f"{AUG}.Normalize", mean=np.array([0, 0, 0]), stds=np.array([1, 1, 1]), image_keys=[("original_image", i) for i in FRAMES]),
        edict(name=f"{AUG}.ConvertToTensor"),
    ], **mappings)


def val_cfg():
    resize, color, mappings = _keys()
    return edict(name="vision_base.utils.builder.Sequential", cfg_list=[
        edict(name=f"{AUG}.ConvertToFloat"),
        edict(name=f"{AUG}.Resize", size=(OUT_H, OUT_W), preserve_aspect_ratio=False),
        edict(name=f"{AUG}.Normalize", mean=np.array([0.485, 0.456, 0.406]), stds=np.array([0.229, 0.224, 0.225]), image_keys=color),
        edict(name=f"{AUG}.Normalize", mean=np.array([0, 0, 0]), stds=np.array([1, 1, 1]), image_keys=[("original_image", i) for i in FRAMES]),
        edict(name=f"{AUG}.ConvertToTensor"),
    ], **mappings)


def summarize(out):
    """Full arrays for two images, the mask, the calibration and the poses; (sum, abs-sum) for every other entry."""
    res = {}
    for k, v in out.items():
        a = np.asarray(v.numpy() if hasattr(v, "numpy") else v)
        name = "_".join(map(str, k)) if isinstance(k, tuple) else str(k)
        if k in (("image", 0), ("original_image", 1), "P2", "patched_mask") or (isinstance(k, tuple) and k[0] == "relative_pose"):
            res["full/" + name] = a
            res["dtype/" + name] = np.array(str(a.dtype))
        else:
            res["sum/" + name] = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
    return res


def fisheye_train_cfg(size=(64, 64)):
    """The augmentation list of configs/kitti360_fisheye_example:131-163 (Resize + mirror + Copy to original_image + colour)."""
    frames = [0, 1, -1]
    image_keys = [("image", i) for i in frames]
    original_keys = [("original_image", i) for i in frames]
    poses = [(("relative_pose", i), 0) for i in frames[1:]]
    return edict(name="vision_base.utils.builder.Sequential", cfg_list=[
        edict(name=f"{AUG}.ConvertToFloat"),
        edict(name=f"{AUG}.Resize", size=size, preserve_aspect_ratio=True, force_pad=True),
        edict(name=f"{AUG}.RandomMirror", mirror_prob=0.5, pose_axis_pairs=poses),
        edict(name=f"{AUG}.Copy", from_keys=image_keys, to_keys=original_keys),
        edict(name="vision_base.utils.builder.Shuffle", cfg_list=[
            edict(name=f"{AUG}.RandomBrightness", distort_prob=1.0),
            edict(name=f"{AUG}.RandomContrast", distort_prob=1.0, lower=0.6, upper=1.4),
            edict(name="vision_base.utils.builder.Sequential", cfg_list=[
                edict(name=f"{AUG}.ConvertColor", transform="HSV"),
                edict(name=f"{AUG}.RandomSaturation", distort_prob=1.0, lower=0.6, upper=1.4),
                edict(name=f"{AUG}.ConvertColor", current="HSV", transform="RGB"),
            ]),
        ]),
        edict(name=f"{AUG}.Normalize", mean=np.array([0.485, 0.456, 0.406]), stds=np.array([0.229, 0.224, 0.225]), image_keys=image_keys),
        edict(name=f"{AUG}.Normalize", mean=np.array([0, 0, 0]), stds=np.array([1, 1, 1]), image_keys=original_keys),
        edict(name=f"{AUG}.ConvertToTensor", image_keys=image_keys + original_keys),
    ], image_keys=image_keys, calib_keys=["P2"], gt_image_keys=["patched_mask"])


def nusc_train_cfg(size=(64, 128)):
    """The augmentation list of configs/nusc_wpose_example:123-151 (aspect-preserving Resize + pad, colour, mirror)."""
    frames = [0, 1, -1]
    image_keys = [("image", i) for i in frames]
    original_keys = [("original_image", i) for i in frames]
    poses = [(("relative_pose", i), 0) for i in frames[1:]]
    return edict(name="vision_base.utils.builder.Sequential", cfg_list=[
        edict(name=f"{AUG}.ConvertToFloat"),
        edict(name=f"{AUG}.Resize", size=size, preserve_aspect_ratio=True, force_pad=True),
        edict(name="vision_base.utils.builder.Shuffle", image_keys=image_keys, cfg_list=[
            edict(name=f"{AUG}.RandomBrightness", distort_prob=1.0),
            edict(name=f"{AUG}.RandomContrast", distort_prob=1.0, lower=0.6, upper=1.4),
            edict(name="vision_base.utils.builder.Sequential", cfg_list=[
                edict(name=f"{AUG}.ConvertColor", transform="HSV"),
                edict(name=f"{AUG}.RandomSaturation", distort_prob=1.0, lower=0.6, upper=1.4),
                edict(name=f"{AUG}.ConvertColor", current="HSV", transform="RGB"),
            ]),
        ]),
        edict(name=f"{AUG}.RandomMirror", mirror_prob=0.5, pose_axis_pairs=poses),
        edict(name=f"{AUG}.Normalize", mean=np.array([0.485, 0.456, 0.406]), stds=np.array([0.229, 0.224, 0.225]), image_keys=image_keys),
        edict(name=f"{AUG}.Normalize", mean=np.array([0, 0, 0]), stds=np.array([1, 1, 1]), image_keys=original_keys),
        edict(name=f"{AUG}.ConvertToTensor"),
    ], image_keys=image_keys + original_keys, calib_keys=["P2"], gt_image_keys=["patched_mask"])


def eval_cfg(size=(OUT_H, OUT_W), preserve_aspect_ratio=False):
    """The single-frame evaluation list of the reference configs (configs/kitti_wpose_example:160-171)."""
    return edict(name="vision_base.utils.builder.Sequential", cfg_list=[
        edict(name=f"{AUG}.ConvertToFloat"),
        edict(name=f"{AUG}.Resize", size=size, preserve_aspect_ratio=preserve_aspect_ratio, force_pad=True),
        edict(name=f"{AUG}.Normalize", mean=np.array([0.485, 0.456, 0.406]), stds=np.array([0.229, 0.224, 0.225])),
        edict(name=f"{AUG}.ConvertToTensor"),
    ], image_keys=[("image", 0)], calib_keys=["P2"])
