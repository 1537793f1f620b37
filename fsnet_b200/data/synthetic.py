"""Synthetic image-triplet dataset with the reference's sample-dict schema (SURVEY.md section 8(b),
mono_dataset.py:179-218): tuple keys ('image', f), ('original_image', f), ('relative_pose', f),
'P2', 'original_P2', fp64 'patched_mask'.  Lets scripts/train.py and bench.py run without KITTI."""
import math

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def make_batch(B, H, W, seed=1234, frame_ids=(0, 1, -1), mask_dtype=torch.float64, fx_scale=0.58, fy_scale=1.92,
               device="cpu"):
    """Seeded smooth-plus-noise triplets, KITTI-like normalised intrinsics, forward-motion poses and a
    patched mask with zero border strips on half of the samples.  (Same recipe as the oracle's
    ``synthetic_batch``; kept separate because product code must not import the oracle.)"""
    g = torch.Generator().manual_seed(seed)
    data = {}
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    base = torch.rand(B, 3, max(H // 8, 2), max(W // 8, 2), generator=g)
    for f in frame_ids:
        lo = base + 0.15 * torch.rand(base.shape, generator=g)
        img = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
        img = (img / 1.15 + 0.05 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
        data[("original_image", f)] = img.contiguous()
        data[("image", f)] = ((img - mean) / std).contiguous()
    P2 = torch.zeros(B, 3, 4)
    P2[:, 0, 0], P2[:, 0, 2] = fx_scale * W, 0.5 * W
    P2[:, 1, 1], P2[:, 1, 2] = fy_scale * H, 0.5 * H
    P2[:, 2, 2] = 1.0
    data["P2"] = P2
    data["original_P2"] = P2.double()
    for f in frame_ids[1:]:
        ang = (torch.rand(B, generator=g) * 2 - 1) * math.radians(1.0)
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, 0, 0], T[:, 0, 2] = torch.cos(ang), torch.sin(ang)
        T[:, 2, 0], T[:, 2, 2] = -torch.sin(ang), torch.cos(ang)
        T[:, 2, 3] = (0.8 + 0.2 * (torch.rand(B, generator=g) * 2 - 1)) * (-1.0 if f > 0 else 1.0)
        T[:, 0, 3] = 0.05 * (torch.rand(B, generator=g) * 2 - 1)
        data[("relative_pose", f)] = T
    mask = torch.ones(B, H, W, dtype=mask_dtype)
    for b in range(0, B, 2):
        wstrip = int(torch.randint(1, max(W // 10, 2), (1,), generator=g))
        if (b // 2) % 2 == 0:
            mask[b, :, :wstrip] = 0
        else:
            mask[b, :, W - wstrip:] = 0
    data["patched_mask"] = mask
    if device != "cpu":
        data = {k: v.to(device) for k, v in data.items()}
    return data


MEI_KITTI360 = dict(xi=2.2134, k1=0.016798, k2=1.6548, gamma=1336.3, u0=716.94, v0=705.76, size=1400.0)


def make_fisheye_batch(B, H, W, seed=1234, frame_ids=(0, 1, -1), mask_dtype=torch.float64, device="cpu"):
    """``make_batch`` with a KITTI-360-like MEI calibration scaled to the crop, lateral (side-looking camera)
    motion and ``calib_meta`` dicts as fisheye_dataset.py:45-58,254 delivers them."""
    data = make_batch(B, H, W, seed, frame_ids, mask_dtype)
    g = torch.Generator().manual_seed(seed + 77)
    m = MEI_KITTI360
    P2 = torch.zeros(B, 3, 4)
    P2[:, 0, 0], P2[:, 1, 1] = m["gamma"] * W / m["size"], m["gamma"] * H / m["size"]
    P2[:, 0, 2], P2[:, 1, 2] = m["u0"] * W / m["size"], m["v0"] * H / m["size"]
    P2[:, 2, 2] = 1.0
    data["P2"], data["original_P2"] = P2, P2.double()
    for f in frame_ids[1:]:
        T = data[("relative_pose", f)]
        T[:, 0, 3] = (0.8 + 0.2 * (torch.rand(B, generator=g) * 2 - 1)) * (-1.0 if f > 0 else 1.0)
        T[:, 2, 3] = 0.1 * (torch.rand(B, generator=g) * 2 - 1)
    if device != "cpu":
        data = {k: v.to(device) for k, v in data.items()}
    data["calib_meta"] = [dict(mirror_parameters=dict(xi=m["xi"]), distortion_parameters=dict(k1=m["k1"], k2=m["k2"]))
                          for _ in range(B)]
    return data


class SyntheticTripletDataset(torch.utils.data.Dataset):
    """``length`` samples of size (height, width); sample i is reproducible from (seed, i)."""

    def __init__(self, length=256, height=192, width=640, frame_idxs=(0, 1, -1), seed=1234, fisheye=False, **kwargs):
        self.length, self.height, self.width = int(length), int(height), int(width)
        self.frame_idxs, self.seed, self.fisheye = tuple(frame_idxs), int(seed), bool(fisheye)

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        make = make_fisheye_batch if self.fisheye else make_batch
        batch = make(1, self.height, self.width, seed=self.seed + int(index), frame_ids=self.frame_idxs)
        return {k: v[0] for k, v in batch.items()}
