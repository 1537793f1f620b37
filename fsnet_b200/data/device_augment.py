"""Training augmentation with the pixel work on the GPU.

The reference augments every sample on the CPU in float32 (``ConvertToFloat -> RandomWarpAffine (KITTI) or Resize + pad
(nuScenes, KITTI-360 fisheye) -> RandomMirror -> Shuffle[RandomBrightness, RandomContrast, HSV/RandomSaturation/RGB] -> Normalize x2 ->
ConvertToTensor``, configs/kitti_wpose_example:123-158, nusc_wpose_example:123-151, kitti360_fisheye_example:131-163) and ships 6 float32 frames per sample to the device.  Here the loader worker only DRAWS the
random parameters -- with the very same augmentation objects, built from the very same config list, so a seeded run makes the
same draws -- and updates the small entries (calibration, poses); the uint8 frames travel as they were decoded (8x fewer bytes
than 2 x float32) and one CUDA kernel (csrc/augment.cu) does warp + mirror + colour chain + normalisation for all frames of
the batch, writing ``('image', f)``, ``('original_image', f)`` and ``patched_mask`` in the layout the model consumes.

    train_dataset.augmentation = edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=<the reference's list>)
    loader  = build_dataloader(dataset, ..., collate_fn=device_augment_collate)
    batches = DevicePrefetcher(loader, device_transform=DeviceAugmentStage(dataset.transform))

Pixel arithmetic follows OpenCV's (fixed-point warp coordinates, float HSV): oracle/augment_oracle.py restates it and is pinned
against cv2 and against the reference pipeline's golden vectors.  One deliberate difference: OpenCV 4.13 warps CV_64F images with
INTER_NEAREST WITHOUT inverting the matrix (the fp64 ``patched_mask`` of the reference lands in the wrong place under that
version); the mask here follows the documented semantics, i.e. what every other depth and older versions do.
"""
from typing import Dict, List

import numpy as np
import torch

from ..utils.builder import Sequential, Shuffle, build
from . import augmentations as A
from .loading import collate_fn

OP_NONE, OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION = 0, 1, 2, 3
PLAN_SIZE = 16           # [0:6] geometry, 6 mirror, 7:10 op codes, 10:13 op values, 13 h0, 14 w0, 15 geometry mode
GEOM_AFFINE, GEOM_RESIZE = 0, 1   # [0:6] = inverse affine (row major)  |  scale_x, scale_y, w_eff, h_eff, -, -


def _invert_affine(M32: np.ndarray) -> np.ndarray:
    """The inversion cv2.warpAffine applies to a forward matrix, in double (see oracle/augment_oracle.py:invert_affine)."""
    M = M32.astype(np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0], M[1, 1] = a11, a22
    M[0, 1] *= -D
    M[1, 0] *= -D
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def _is_saturation_block(obj) -> bool:
    c = getattr(obj, "children", None)
    return (isinstance(obj, Sequential) and c is not None and len(c) == 3 and isinstance(c[0], A.ConvertColor)
            and isinstance(c[1], A.RandomSaturation) and isinstance(c[2], A.ConvertColor)
            and (c[0].current, c[0].transform, c[2].current, c[2].transform) == ("RGB", "HSV", "HSV", "RGB"))


class DeviceAugmentation(object):
    """Host half: called by the dataset reader on the raw sample (uint8 frames, P2, poses, mask) in the loader worker."""

    def __init__(self, pipeline: Dict, **kwargs):
        self.pipe = build(**pipeline)
        if not isinstance(self.pipe, Sequential):
            raise NotImplementedError("DeviceAugmentation expects the reference's Sequential augmentation list")
        self.geom = self.mirror = self.norm = None
        self.steps: List = []            # ("geom",), ("mirror",), ("colour", op) / ("shuffle", Shuffle) in list order
        for child in self.pipe.children:
            if isinstance(child, (A.ConvertToFloat, A.ConvertToTensor)):
                continue
            if isinstance(child, (A.RandomWarpAffine, A.Resize)) and self.geom is None and not self.steps:
                self.geom = child
                self.steps.append(("geom",))
            elif isinstance(child, A.RandomMirror) and self.mirror is None:
                self.mirror = child
                self.steps.append(("mirror",))
            elif isinstance(child, Shuffle) and all(self._colour_code(c) for c in child.children):
                self.steps.append(("shuffle", child))
            elif self._colour_code(child):
                self.steps.append(("colour", child))
            elif isinstance(child, A.Copy) and all(isinstance(a, tuple) and isinstance(b, tuple) and a[0] == "image"
                                                   and b[0] == "original_image" and a[1] == b[1]
                                                   for a, b in zip(child.from_keys, child.to_keys)) and not any(
                                                       st[0] in ("shuffle", "colour") for st in self.steps):
                continue                                       # original_image := the geometry-only frame: what the kernel writes
            elif isinstance(child, A.Normalize):
                if np.all(child.mean == 0) and np.all(child.stds == 1):
                    continue                                   # the 'original_image' entries: plain / 255
                if self.norm is not None:
                    raise NotImplementedError("DeviceAugmentation: more than one mean/std normalisation in the list")
                self.norm = child
            else:
                raise NotImplementedError(f"DeviceAugmentation: {type(child).__name__} has no device implementation here (supported "
                                          "lists: RandomWarpAffine or Resize first, then RandomMirror / colour jitter / Copy / Normalize)")
        if self.geom is None or self.norm is None:
            raise NotImplementedError("DeviceAugmentation needs a RandomWarpAffine or Resize first and a mean/std Normalize in the list")
        self.is_warp = isinstance(self.geom, A.RandomWarpAffine)
        if self.is_warp and self.geom.border_mode != 0:
            raise NotImplementedError("DeviceAugmentation: only cv2.BORDER_CONSTANT warps")
        if not self.is_warp and len(self.geom.size) != 2:
            raise NotImplementedError("DeviceAugmentation: Resize(size=(h, w)) expected")
        n_ops = sum(len(s[1].children) if s[0] == "shuffle" else 1 for s in self.steps if s[0] in ("shuffle", "colour"))
        if n_ops > 3:
            raise NotImplementedError("DeviceAugmentation: at most three colour operations")
        self.frames = [k[1] for k in self.norm.image_keys if isinstance(k, tuple) and k[0] == "image"]
        if not self.frames:
            raise NotImplementedError("DeviceAugmentation: Normalize(image_keys=[('image', f), ...]) expected")
        if self.is_warp:
            self.output_h, self.output_w = self.geom.output_h, self.geom.output_w
        else:
            self.output_h, self.output_w = int(self.geom.size[0]), int(self.geom.size[1])
        self.mean, self.std = self.norm.mean.astype(np.float32), self.norm.stds.astype(np.float32)

    @staticmethod
    def _colour_code(obj) -> int:
        if isinstance(obj, A.RandomBrightness):
            return OP_BRIGHTNESS
        if isinstance(obj, A.RandomContrast):
            return OP_CONTRAST
        if _is_saturation_block(obj):
            return OP_SATURATION
        return OP_NONE

    def _draw_colour(self, op, codes, values):
        code = self._colour_code(op)
        value = (op.children[1] if code == OP_SATURATION else op).draw()
        if code == OP_SATURATION:
            # the HSV round trip happens whether or not the factor is drawn (it is not bit-neutral): NaN = no factor
            codes.append(code)
            values.append(np.nan if value is None else value)
        elif value is not None:
            codes.append(code)
            values.append(value)

    def __call__(self, data: Dict) -> Dict:
        first = data[("image", self.frames[0])]
        h0, w0 = first.shape[:2]
        plan = np.zeros(PLAN_SIZE, dtype=np.float64)
        codes, values = [], []
        for step in self.steps:
            if step[0] == "geom" and self.is_warp:
                s, shift_w, shift_h = self.geom.draw(h0, w0)
                M = np.array([[s, 0, shift_w], [0, s, shift_h]], dtype=np.float32)
                plan[0:6] = _invert_affine(M).reshape(-1)
                self.geom.warp_calibration(data, s, shift_w, shift_h)
            elif step[0] == "geom":
                h, w, _, scale_yx = self.geom.geometry(h0, w0)
                data[("image_resize", "original_shape")] = np.array([h0, w0]).astype(int)
                data[("image_resize", "effective_size")] = np.array([h, w]).astype(int)
                plan[0:4] = 1.0 / (w / w0), 1.0 / (h / h0), w, h          # cv2.resize: scale = 1 / (dsize / ssize), in double
                plan[15] = GEOM_RESIZE
                self.geom.resize_calibration(data, scale_yx)
            elif step[0] == "mirror":
                if self.mirror.draw():
                    plan[6] = 1.0
                    self.mirror.mirror_entries(data, self.output_w)
            elif step[0] == "shuffle":
                for i in step[1].draw():
                    self._draw_colour(step[1].children[i], codes, values)
            else:
                self._draw_colour(step[1], codes, values)
        plan[7:7 + len(codes)] = codes
        plan[10:10 + len(values)] = values
        plan[13], plan[14] = h0, w0
        frames = np.stack([np.ascontiguousarray(data.pop(("image", f))) for f in self.frames])
        if frames.dtype != np.uint8:
            raise NotImplementedError("DeviceAugmentation: the frames must arrive as decoded uint8 images")
        for f in self.frames:
            data.pop(("original_image", f), None)          # the same pixels before augmentation: rebuilt on the device
        mask = data.pop("patched_mask", None)
        if mask is not None:
            mask_u8 = np.asarray(mask).astype(np.uint8)
            if not np.array_equal(mask_u8, mask):
                raise NotImplementedError("DeviceAugmentation: patched_mask must hold 0 / 1")
            data["mask_u8"] = mask_u8
            data["mask_dtype"] = str(np.asarray(mask).dtype)      # the list keeps the mask's dtype (fp64 ones / a uint8 validity image)
        data["frames_u8"] = frames
        data["aug_plan"] = plan
        for key in self.geom.calib_keys:                   # ConvertToTensor (augmentations.py:62-89)
            data[key] = torch.tensor(data[key], dtype=torch.float32).contiguous()
        return data


def device_augment_collate(batch: List[Dict]) -> Dict:
    """collate_fn for samples of DeviceAugmentation: frames / masks of different sizes (KITTI drives differ by a few pixels)
    are zero-padded to the largest of the batch; the true size is in the plan."""
    h = max(b["frames_u8"].shape[1] for b in batch)
    w = max(b["frames_u8"].shape[2] for b in batch)
    for b in batch:
        f = b["frames_u8"]
        if f.shape[1] != h or f.shape[2] != w:
            b["frames_u8"] = np.pad(f, ((0, 0), (0, h - f.shape[1]), (0, w - f.shape[2]), (0, 0)))
            if "mask_u8" in b:
                m = b["mask_u8"]
                b["mask_u8"] = np.pad(m, ((0, h - m.shape[0]), (0, w - m.shape[1])))
    return collate_fn(batch)


class DeviceAugmentStage(object):
    """Device half: turns ``frames_u8`` / ``mask_u8`` / ``aug_plan`` of an uploaded batch into the model's inputs with one
    launch of ``fsnet_augment_frames``.  CUDA only."""

    def __init__(self, augmentation: DeviceAugmentation, ring: int = 2):
        self.frames = list(augmentation.frames)
        self.output_h, self.output_w = augmentation.output_h, augmentation.output_w
        self.mean_std = torch.tensor(np.concatenate([augmentation.mean, augmentation.std]), dtype=torch.float32)
        # output buffers are kept and used round-robin, `ring` = the DevicePrefetcher's slots (depth + 1): a set is rewritten only
        # after the prefetcher has seen the consumer release the slot it belongs to (no allocator traffic on the upload stream)
        self._ring, self._sets, self._turn = max(int(ring), 1), {}, 0

    def _buffers(self, key, dev, F, B, H, W, with_mask):
        sets = self._sets.setdefault(key, [])
        if len(sets) < self._ring:
            sets.append((torch.empty(F, B, 3, H, W, device=dev, dtype=torch.float32), torch.empty(F, B, 3, H, W, device=dev, dtype=torch.float32),
                         torch.empty(B, H, W, device=dev, dtype=torch.float64) if with_mask else None))
            return sets[-1]
        self._turn = (self._turn + 1) % self._ring
        return sets[self._turn]

    def __call__(self, batch: Dict) -> Dict:
        from .. import _lib
        frames = batch.pop("frames_u8")
        plan = batch.pop("aug_plan")
        mask = batch.pop("mask_u8", None)
        if not frames.is_cuda:
            raise _lib.FsnetError("DeviceAugmentStage runs on an uploaded batch (there is no CPU path)")
        B, F, H0, W0, _ = frames.shape
        dev = frames.device
        if self.mean_std.device != dev:
            self.mean_std = self.mean_std.to(dev)
        H, W = self.output_h, self.output_w
        image, original, mask_out = self._buffers((B, F, H, W, mask is not None, str(dev)), dev, F, B, H, W, mask is not None)
        _lib.call("fsnet_augment_frames", frames.contiguous(), None if mask is None else mask.contiguous(), plan.double().contiguous(),
                  B, F, H0, W0, H, W, self.mean_std, image, original, mask_out)
        for i, f in enumerate(self.frames):
            batch[("image", f)] = image[i]
            batch[("original_image", f)] = original[i]
        dtypes = batch.pop("mask_dtype", None)
        if mask_out is not None:
            keep_integer = bool(dtypes) and all(d == "uint8" for d in dtypes)
            batch["patched_mask"] = mask_out.to(torch.uint8) if keep_integer else mask_out
        return batch


def find_device_stage(dataset):
    """The DeviceAugmentStage of a dataset whose reader (or every reader of a ConcatDataset) augments through
    DeviceAugmentation; None when the dataset augments on the host.  Mixed datasets are refused."""
    readers = list(getattr(dataset, "children", None) or [dataset])
    augs = [getattr(r, "transform", None) for r in readers]
    flagged = [isinstance(a, DeviceAugmentation) for a in augs]
    if not any(flagged):
        return None
    if not all(flagged):
        raise NotImplementedError("every reader of the dataset must use DeviceAugmentation (or none)")
    first = augs[0]
    for a in augs[1:]:
        if (a.frames, a.output_h, a.output_w) != (first.frames, first.output_h, first.output_w) or not (
                np.array_equal(a.mean, first.mean) and np.array_equal(a.std, first.std)):
            raise NotImplementedError("the readers' DeviceAugmentation settings differ (output size / frames / normalisation)")
    return DeviceAugmentStage(first)
