"""Host-side sample augmentations with the reference's class names, keyword surface and random streams
(vision_base/data/augmentations/augmentations.py; the ones the four in-scope configs build).

These run in the DataLoader workers on numpy arrays exactly like the reference (OpenCV + numpy): they are the *caller*
side of the hot path.  A transform takes and returns the sample dict; which entries it touches is given by
``image_keys`` / ``gt_image_keys`` / ``calib_keys`` (``Sequential(..., **key_mappings)`` hands the same lists to every
child).  Random draws follow the reference's generators call for call (``default_rng`` seeded from ``np.random`` when no
``random_seed`` is given; ``RandomMirror`` draws from ``np.random`` directly), so a seeded pipeline reproduces the
reference's samples bit for bit -- pinned by tests/golden/aug_*.npz.
"""
import cv2
import numpy as np
import torch


def _rng(random_seed):
    # augmentations.py:216,459: one generator per transform, seeded from the global numpy stream unless fixed
    return np.random.default_rng(random_seed if random_seed is not None else np.random.randint(0, 2 ** 32))


def flip_relative_pose(pose: np.ndarray, axis_num=0) -> np.ndarray:
    """Pose of the mirrored world (vision_base/data/augmentations/utils.py:4-22).  The reference negates the two Euler
    angles ('xyz') that do not belong to ``axis_num`` and the translation along it; that is the conjugation F T F with
    the reflection F = diag(-1 at axis_num), written here in closed form."""
    f = np.ones(3, dtype=np.float64)
    f[axis_num] = -1.0
    out = np.eye(4, dtype=np.float32)
    out[:3, :3] = (f[:, None] * pose[:3, :3].astype(np.float64) * f[None, :]).astype(np.float32)
    out[:3, 3] = (f * pose[:3, 3].astype(np.float64)).astype(np.float32)
    return out


class ConvertToFloat(object):
    """uint8 -> float32 for the image entries (augmentations.py:50-60)."""

    def __init__(self, image_keys=("image",), **kwargs):
        self.image_keys = list(image_keys)

    def __call__(self, data):
        for key in self.image_keys:
            data[key] = data[key].astype(np.float32)
        return data


class ConvertToTensor(object):
    """HWC arrays -> CHW float32 tensors; 2-D ground-truth maps keep their dtype (the fp64 ``patched_mask`` of SURVEY.md
    App. C-3 comes from here); calibration / lidar entries -> float32 tensors (augmentations.py:62-89)."""

    def __init__(self, image_keys=("image",), gt_image_keys=(), calib_keys=(), lidar_keys=(), **kwargs):
        self.image_keys, self.gt_image_keys = list(image_keys), list(gt_image_keys)
        self.calib_keys, self.lidar_keys = list(calib_keys), list(lidar_keys)

    def __call__(self, data):
        for key in self.image_keys + self.gt_image_keys:
            arr = data[key]
            if arr.ndim == 3:
                data[key] = torch.tensor(arr.transpose(2, 0, 1), dtype=torch.float32).contiguous()
            else:
                data[key] = torch.tensor(arr).contiguous()
        for key in self.calib_keys + self.lidar_keys:
            data[key] = torch.tensor(data[key], dtype=torch.float32).contiguous()
        return data


class Normalize(object):
    """(image / 255 - mean) / std per channel; mean / std are tiled when the entry stacks several RGB frames
    (augmentations.py:91-109)."""

    def __init__(self, mean, stds, image_keys=("image",), **kwargs):
        self.mean = np.array(mean, dtype=np.float32)
        self.stds = np.array(stds, dtype=np.float32)
        self.image_keys = list(image_keys)

    def __call__(self, data):
        for key in self.image_keys:
            img = data[key].astype(np.float32)
            reps = img.shape[2] // self.mean.shape[0]
            img /= 255.0
            img -= np.tile(self.mean, reps)
            img /= np.tile(self.stds, img.shape[2] // self.stds.shape[0])
            data[key] = img
        return data


class Resize(object):
    """Resize to ``size`` = (h, w).  ``preserve_aspect_ratio``: one scale factor, then zero-pad (``force_pad``) or crop /
    pad the width; otherwise anisotropic.  Ground-truth maps use nearest neighbour; rows 0 / 1 of the calibration
    matrices are scaled by the x / y factors; the original and effective sizes are recorded (augmentations.py:112-198)."""

    def __init__(self, size, preserve_aspect_ratio=True, force_pad=True, image_keys=("image",), calib_keys=(), gt_image_keys=(), **kwargs):
        self.size, self.preserve_aspect_ratio, self.force_pad = size, preserve_aspect_ratio, force_pad
        self.image_keys, self.calib_keys, self.gt_image_keys = list(image_keys), list(calib_keys), list(gt_image_keys)

    def geometry(self, h0, w0):
        """(h, w) of the resized frame, how it is brought to ``size`` afterwards ('none' | 'pad_0' | 'pad_1' | 'crop_1') and the
        (y, x) factors applied to the calibration."""
        mode = "none"
        if self.preserve_aspect_ratio:
            fy, fx = self.size[0] / h0, self.size[1] / w0          # the reference calls these scale_factor_x / _y
            if self.force_pad:
                f = min(fy, fx)
                mode = "pad_0" if fy > fx else "pad_1"
            else:
                f = fy
                mode = "crop_1" if fy > fx else "pad_1"
            return int(np.round(h0 * f)), int(np.round(w0 * f)), mode, (f, f)
        return self.size[0], self.size[1], mode, (self.size[0] / h0, self.size[1] / w0)

    def resize_calibration(self, data, scale_yx):
        for key in self.calib_keys:
            P = data[key]
            P[0, :] = P[0, :] * scale_yx[1]
            P[1, :] = P[1, :] * scale_yx[0]
            data[key] = P

    def __call__(self, data):
        h0, w0 = data[self.image_keys[0]].shape[:2]
        data[("image_resize", "original_shape")] = np.array([h0, w0]).astype(int)
        h, w, mode, scale_yx = self.geometry(h0, w0)
        data[("image_resize", "effective_size")] = np.array([h, w]).astype(int)
        for key in self.image_keys:
            data[key] = cv2.resize(data[key], (w, h))
        for key in self.gt_image_keys:
            data[key] = cv2.resize(data[key], (w, h), interpolation=cv2.INTER_NEAREST)
        if len(self.size) > 1:
            for key in self.image_keys + self.gt_image_keys:
                img = data[key]
                if mode == "crop_1":
                    data[key] = img[:, 0:self.size[1]]
                elif mode in ("pad_1", "pad_0"):
                    pad = [(0, 0)] * img.ndim
                    if mode == "pad_1":
                        pad[1] = (0, self.size[1] - img.shape[1])
                    else:
                        pad[0] = (0, self.size[0] - img.shape[0])
                    data[key] = np.pad(img, pad, "constant")
        self.resize_calibration(data, scale_yx)
        return data


class RandomSaturation(object):
    """Scales the S channel of an HSV image by U(lower, upper) with probability ``distort_prob`` (augmentations.py:200-226)."""

    def __init__(self, distort_prob, lower=0.5, upper=1.5, image_keys=("image",), random_seed=None, **kwargs):
        assert upper >= lower >= 0, "saturation bounds must satisfy 0 <= lower <= upper"
        self.distort_prob, self.lower, self.upper = distort_prob, lower, upper
        self.image_keys = list(image_keys)
        self.rng = _rng(random_seed)

    def draw(self):
        """The factor of this call, or None (same generator calls as the reference, so seeded runs agree)."""
        return self.rng.uniform(self.lower, self.upper) if self.rng.random() <= self.distort_prob else None

    def __call__(self, data):
        ratio = self.draw()
        if ratio is not None:
            for key in self.image_keys:
                data[key][:, :, 1] *= ratio
        return data


class RandomContrast(object):
    """image *= U(lower, upper) with probability ``distort_prob`` (augmentations.py:545-570)."""

    def __init__(self, distort_prob, lower=0.5, upper=1.5, image_keys=("image",), random_seed=None, **kwargs):
        assert upper >= lower >= 0, "contrast bounds must satisfy 0 <= lower <= upper"
        self.distort_prob, self.lower, self.upper = distort_prob, lower, upper
        self.image_keys = list(image_keys)
        self.rng = _rng(random_seed)

    def draw(self):
        return self.rng.uniform(self.lower, self.upper) if self.rng.random() <= self.distort_prob else None

    def __call__(self, data):
        alpha = self.draw()
        if alpha is not None:
            for key in self.image_keys:
                data[key] = data[key] * alpha
        return data


class RandomBrightness(object):
    """image += U(-delta, delta) with probability ``distort_prob`` (augmentations.py:572-592)."""

    def __init__(self, distort_prob, delta=32, image_keys=("image",), random_seed=None, **kwargs):
        assert 0.0 <= delta <= 255.0
        self.distort_prob, self.delta = distort_prob, delta
        self.image_keys = list(image_keys)
        self.rng = _rng(random_seed)

    def draw(self):
        return self.rng.uniform(-self.delta, self.delta) if self.rng.random() <= self.distort_prob else None

    def __call__(self, data):
        delta = self.draw()
        if delta is not None:
            for key in self.image_keys:
                data[key] = data[key] + delta
        return data


class ConvertColor(object):
    """cv2.cvtColor between colour spaces, e.g. RGB <-> HSV (augmentations.py:527-543)."""

    def __init__(self, current="RGB", transform="HSV", image_keys=("image",), **kwargs):
        self.current, self.transform = current, transform
        self.image_keys = list(image_keys)
        self.convertor = getattr(cv2, f"COLOR_{current}2{transform}")

    def __call__(self, data):
        for key in self.image_keys:
            data[key] = cv2.cvtColor(data[key], self.convertor)
        return data


class RandomMirror(object):
    """Horizontal flip with probability ``mirror_prob`` (drawn from the global numpy stream): images and ground-truth maps
    are reversed along x, P[0,3] := -P[0,3] and P[0,2] := W - P[0,2] - 1, relative poses are mirrored about their axis, lidar
    x is negated, stereo pairs swap sides (augmentations.py:377-434)."""

    def __init__(self, mirror_prob, image_keys=("image",), calib_keys=(), gt_image_keys=(), object_keys=(), lidar_keys=(),
                 pose_axis_pairs=(), is_switch_left_right=True, stereo_image_key_pairs=(), stereo_calib_key_pairs=(), **kwargs):
        self.mirror_prob = mirror_prob
        self.image_keys, self.calib_keys, self.gt_image_keys = list(image_keys), list(calib_keys), list(gt_image_keys)
        self.object_keys, self.lidar_keys = list(object_keys), list(lidar_keys)
        self.pose_axis_pairs = list(pose_axis_pairs)
        self.is_switch_lr = is_switch_left_right
        self.stereo_pairs = list(stereo_image_key_pairs) + list(stereo_calib_key_pairs)

    def draw(self) -> bool:
        return bool(np.random.rand() <= self.mirror_prob)

    def mirror_entries(self, data, width):
        """Everything a flip changes apart from the pixel arrays (calibration, objects, lidar, poses, stereo sides)."""
        for key in self.calib_keys:
            P = data[key]
            P[0, 3] = -P[0, 3]
            P[0, 2] = width - P[0, 2] - 1
            data[key] = P
        for key in self.object_keys:
            data[key].flip_objects()
        for key in self.lidar_keys:
            data[key] = -data[key][..., 0]
        for key, axis_num in self.pose_axis_pairs:
            data[key] = flip_relative_pose(data[key], axis_num)
        if self.is_switch_lr:
            for left, right in self.stereo_pairs:
                data[left], data[right] = data[right], data[left]

    def __call__(self, data):
        width = data[self.image_keys[0]].shape[1]
        if self.draw():
            for key in self.image_keys + self.gt_image_keys:
                data[key] = np.ascontiguousarray(data[key][:, ::-1])
            self.mirror_entries(data, width)
        return data


class RandomWarpAffine(object):
    """Random zoom (scale ~ U(lower, upper) of the longer side) about a random centre, warped straight to the output size
    with one ``cv2.warpAffine`` (bilinear for images, nearest for ground-truth maps, constant border -- the zero strips
    that end up in ``patched_mask``); calibration rows 0 / 1 are scaled and shifted accordingly (augmentations.py:436-498)."""

    def __init__(self, scale_lower=0.6, scale_upper=1.4, shift_border=128, output_w=1280, output_h=384, image_keys=("image",),
                 gt_image_keys=(), calib_keys=(), border_mode=cv2.BORDER_CONSTANT, random_seed=None, **kwargs):
        self.scale_lower, self.scale_upper, self.shift_border = scale_lower, scale_upper, shift_border
        self.output_w, self.output_h = output_w, output_h
        self.image_keys, self.gt_image_keys, self.calib_keys = list(image_keys), list(gt_image_keys), list(calib_keys)
        self.border_mode = border_mode
        self.rng = _rng(random_seed)

    def draw(self, height, width):
        """(s, shift_w, shift_h) of this call: output = s * input + shift."""
        scale = max(height, width) * self.rng.uniform(self.scale_lower, self.scale_upper)
        center_w = self.rng.integers(low=self.shift_border, high=width - self.shift_border)
        center_h = self.rng.integers(low=self.shift_border, high=height - self.shift_border)
        s = max(self.output_w, self.output_h) / scale
        return s, self.output_w / 2 - center_w * s, self.output_h / 2 - center_h * s

    def warp_calibration(self, data, s, shift_w, shift_h):
        for key in self.calib_keys:
            P = data[key]
            P[0:2, :] *= s
            P[0, 2] = P[0, 2] + shift_w
            P[0, 3] = P[0, 3] + shift_w * P[2, 3]
            P[1, 2] = P[1, 2] + shift_h
            P[1, 3] = P[1, 3] + shift_h * P[2, 3]
            data[key] = P

    def __call__(self, data):
        height, width = data[self.image_keys[0]].shape[:2]
        s, shift_w, shift_h = self.draw(height, width)
        M = np.array([[s, 0, shift_w], [0, s, shift_h]], dtype=np.float32)
        size = (self.output_w, self.output_h)
        for key in self.image_keys:
            data[key] = cv2.warpAffine(data[key], M, size, flags=cv2.INTER_LINEAR, borderMode=self.border_mode)
        for key in self.gt_image_keys:
            data[key] = cv2.warpAffine(data[key], M, size, flags=cv2.INTER_NEAREST, borderMode=self.border_mode)
        self.warp_calibration(data, s, shift_w, shift_h)
        return data


class Copy(object):
    """data[to_key] = copy of data[from_key] (augmentations.py:668-680)."""

    def __init__(self, from_keys, to_keys, **kwargs):
        self.from_keys, self.to_keys = list(from_keys), list(to_keys)

    def __call__(self, data):
        for src, dst in zip(self.from_keys, self.to_keys):
            data[dst] = data[src].copy()
        return data


class EmptyAug(object):
    """Identity (augmentations.py:20-28)."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, data):
        return data
