"""KITTI raw readers with the reference's class names, constructor keywords and sample schema
(monodepth/data/datasets/mono_dataset.py:110-345, utils.py:20-58): the callers' side of the hot path.

Directory layout expected under ``raw_path`` (KITTI raw + the poses exported by the MATLAB devkit):
    <date>/calib_cam_to_cam.txt, calib_velo_to_cam.txt, calib_imu_to_velo.txt
    <date>/<drive>_sync/image_02|image_03/data/%010d.png,  <date>/<drive>_sync/oxts/pose.mat  ('pose_mat' [N,4,4] imu->world)
Split files list ``<date>/<drive>_sync <frame> <l|r>`` per line.  A sample is a dict with tuple keys: ('image', f) and
('original_image', f) uint8 HxWx3 for f in frame_idxs, 'P2' / 'original_P2' 3x4 of the chosen camera, ('relative_pose', f)
4x4 float32 camera-frame motion target -> frame f, an all-ones fp64 'patched_mask'; the configured augmentation
(``build(**augmentation)``) then turns it into tensors.
"""
import os
from copy import deepcopy
from typing import List

import cv2
import numpy as np
import torch.utils.data

from ..utils.builder import build


def read_image(path: str) -> np.ndarray:
    """RGB uint8 [H,W,3]."""
    from PIL import Image
    return np.array(Image.open(path, "r"))


def read_depth(path: str) -> np.ndarray:
    """KITTI 16-bit depth png -> metres (float32)."""
    return np.array(cv2.imread(path, -1) / 256.0, dtype=np.float32)


def read_pose_mat(path: str) -> np.ndarray:
    import scipy.io as sio
    return sio.loadmat(path)["pose_mat"]


def cam_relative_pose(T_imu2world_0, T_imu2world_1, T_imu2vel, T_vel2cam):
    """Motion of the camera frame between two IMU poses (utils.py:60-61): cam <- vel <- imu_1 <- world <- imu_0 <- vel <- cam."""
    to_cam = T_vel2cam @ T_imu2vel
    return to_cam @ np.linalg.inv(T_imu2world_1) @ T_imu2world_0 @ np.linalg.inv(T_imu2vel) @ np.linalg.inv(T_vel2cam)


def _numbers_after(line: str, count: int, first: int = 1):
    parts = line.split(" ")
    return np.array([float(x) for x in parts[first:first + count]])


def read_P23_from_sequence(path):
    """P_rect_02 / P_rect_03 (3x4) of calib_cam_to_cam.txt."""
    P = {}
    with open(path) as f:
        for line in f:
            for tag in ("P_rect_02", "P_rect_03"):
                if line.startswith(tag):
                    P[tag] = _numbers_after(line, 12).reshape(3, 4)
    assert "P_rect_02" in P, f"can not find P2 in file {path}"
    assert "P_rect_03" in P, f"can not find P3 in file {path}"
    return P["P_rect_02"], P["P_rect_03"]


def _read_rigid(path, r_tag, t_tag):
    R = t = None
    with open(path) as f:
        for line in f:
            if line.startswith(r_tag):
                R = _numbers_after(line, 9).reshape(3, 3)
            if line.startswith(t_tag):
                t = _numbers_after(line, 3).reshape(3, 1)
    assert R is not None and t is not None, f"can not find R / T in file {path}"
    T = np.eye(4)
    T[:3, :3], T[:3, 3:4] = R, t
    return T


def read_T_from_sequence(path):
    """velodyne -> camera 4x4 of calib_velo_to_cam.txt (lines 'R:' and 'T:')."""
    return _read_rigid(path, "R:", "T:")


def read_imu2velo(path):
    """imu -> velodyne 4x4 of calib_imu_to_velo.txt (any line starting with R / T; the later one wins, as in the reference)."""
    return _read_rigid(path, "R", "T")


def read_split_file(path: str):
    out = []
    with open(path) as f:
        for line in f:
            parts = line.strip().split()
            if not parts:
                continue
            out.append(dict(folder=parts[0], index=int(parts[1]), side=parts[2], datetime=parts[0].split("/")[0]))
    return out


def _calibrations(raw_path):
    meta = {}
    for date in os.listdir(raw_path):
        if not os.path.isdir(os.path.join(raw_path, date)):
            continue
        P2, P3 = read_P23_from_sequence(os.path.join(raw_path, date, "calib_cam_to_cam.txt"))
        meta[date] = dict(P2=P2, P3=P3, T_vel2cam=read_T_from_sequence(os.path.join(raw_path, date, "calib_velo_to_cam.txt")),
                          T_imu2vel=read_imu2velo(os.path.join(raw_path, date, "calib_imu_to_velo.txt")))
    return meta


class _KittiBase(torch.utils.data.Dataset):
    CAMERA = {"l": "image_02", "r": "image_03"}

    def __len__(self):
        return len(self.imdb)

    def get_color(self, folder, frame_index, side):
        return read_image(os.path.join(self.raw_path, folder, self.CAMERA[side], "data", "%010d.png" % frame_index))

    def _relative(self, date, poses, k):
        m = self.meta_dict[date]
        return cam_relative_pose(poses[0], poses[k], m["T_imu2vel"], m["T_vel2cam"]).astype(np.float32)


class KittiDepthMonoDataset(_KittiBase):
    """Training triplets (mono_dataset.py:110-247).  Keywords: raw_path, split_file, frame_idxs, augmentation, depth_path,
    is_filter_static (drops samples whose camera moves less than 3 cm to a neighbour), is_motion_mask / motion_mask_path,
    is_precompute_flow / flow_path."""

    def __init__(self, **data_cfg):
        super().__init__()
        cfg = data_cfg
        self.raw_path = cfg["raw_path"]
        self.depth_path = cfg.get("depth_path")
        self.frame_idxs = list(cfg["frame_idxs"])
        self.imdb = read_split_file(cfg["split_file"])
        self.meta_dict = _calibrations(self.raw_path)
        self.pose_dict = {folder: read_pose_mat(os.path.join(self.raw_path, folder, "oxts", "pose.mat"))
                          for folder in {o["folder"] for o in self.imdb}}
        self.is_motion_mask = cfg.get("is_motion_mask", False)
        self.is_precompute_flow = cfg.get("is_precompute_flow", False)
        self.precompute_path = cfg.get("motion_mask_path", "")
        self.flow_path = cfg.get("flow_path", "")
        self.is_filter_static = cfg.get("is_filter_static", True)
        if self.is_filter_static:
            self.imdb = self._filter_static_indexes()
        self.transform = build(**cfg["augmentation"])

    def get_pose(self, folder, frame_indexes: List[int], *args, **kwargs):
        return self.pose_dict[folder][frame_indexes, :, :]

    def _filter_static_indexes(self):
        print(f"Start Filtering Static indexes, original length {len(self)}")
        keep = []
        for obj in self.imdb:
            poses = self.get_pose(obj["folder"], [obj["index"] + i for i in self.frame_idxs])
            moves = [np.linalg.norm(self._relative(obj["datetime"], poses, k + 1)[:3, 3]) for k in range(len(self.frame_idxs) - 1)]
            if all(m >= 0.03 for m in moves):
                keep.append(obj)
        print(f"Finished filtering static indexes, find dynamic instances {len(keep)}")
        return keep

    def get_depth(self, folder, frame_index, side):
        return read_depth(os.path.join(self.depth_path, folder.split("/")[1], "proj_depth", "groundtruth", self.CAMERA[side],
                                       "%010d.png" % frame_index))

    def get_motion_mask(self, i):
        return cv2.imread(os.path.join(self.precompute_path, f"{i:08d}.png"), cv2.IMREAD_UNCHANGED)

    def get_flow(self, i):
        raw = cv2.imread(os.path.join(self.flow_path, f"{i:08d}.png"), cv2.IMREAD_UNCHANGED)[:, :, 0:2]
        return (raw.astype(np.float32) - 2 ** 15) / 64.0

    def __getitem__(self, i):
        obj = self.imdb[i]
        folder, index, side, date = obj["folder"], obj["index"], obj["side"], obj["datetime"]
        data = {}
        for f in self.frame_idxs:
            data[("image", f)] = self.get_color(folder, index + f, side)
            data[("original_image", f)] = data[("image", f)].copy()
        h, w = data[("image", 0)].shape[:2]
        data["patched_mask"] = np.ones([h, w])
        if self.is_motion_mask:
            data["motion_mask"] = self.get_motion_mask(i)
        if self.is_precompute_flow:
            data["flow"] = self.get_flow(i)
        poses = self.get_pose(folder, [index + f for f in self.frame_idxs])
        for k, f in enumerate(self.frame_idxs[1:]):
            data[("relative_pose", f)] = self._relative(date, poses, k + 1)
        data["P2"] = self.meta_dict[date][{"l": "P2", "r": "P3"}[side]]
        data["original_P2"] = data["P2"].copy()
        if self.depth_path is not None:
            data[("sparse_depth", 0)] = self.get_depth(folder, index, side)
        return self.transform(deepcopy(data))


class KittiDepthMonoEigenTestDataset(_KittiBase):
    """Evaluation samples (mono_dataset.py:250-345): the target frame, its predecessor (the frame itself at index 0), the
    pose to the predecessor; ground-truth depth from ``<raw_path>/<folder>/depth`` when ``depth_path`` is given."""

    def __init__(self, **data_cfg):
        super().__init__()
        self.raw_path = data_cfg["raw_path"]
        self.depth_path = data_cfg.get("depth_path")
        self.imdb = read_split_file(data_cfg["split_file"])
        self.meta_dict = _calibrations(self.raw_path)
        self.transform = build(**data_cfg["augmentation"])

    def get_pose(self, folder, frame_indexes: List[int], *args, **kwargs):
        return read_pose_mat(os.path.join(self.raw_path, folder, "oxts", "pose.mat"))[frame_indexes, :, :]

    def get_depth(self, folder, frame_index, side):
        return read_depth(os.path.join(self.raw_path, folder, "depth", "%010d.png" % frame_index))

    def __getitem__(self, i):
        obj = self.imdb[i]
        folder, index, side, date = obj["folder"], obj["index"], obj["side"], obj["datetime"]
        data = {("image", 0): self.get_color(folder, index, side)}
        data[("image", -1)] = self.get_color(folder, index - 1 if index > 0 else index, side)
        data[("original_image", 0)] = data[("image", 0)].copy()
        data["P2"] = self.meta_dict[date][{"l": "P2", "r": "P3"}[side]]
        data["original_P2"] = data["P2"].copy()
        poses = self.get_pose(folder, [index, index - 1])
        data[("relative_pose", -1)] = self._relative(date, poses, 1)
        if self.depth_path is not None:
            data[("sparse_depth", 0)] = self.get_depth(folder, index, side)
        return self.transform(deepcopy(data))
