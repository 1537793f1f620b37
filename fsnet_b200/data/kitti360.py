"""KITTI-360 readers with the reference's class names, keywords and sample schema: the perspective cameras
(monodepth/data/datasets/kitti360_dataset.py:13-220, ``KITTI360MonoDataset``) and the fisheye cameras
(monodepth/data/datasets/fisheye_dataset.py:17-262, ``KITTI360FisheyeDataset`` -- the caller of the MEI-camera loss head).

Layout under ``raw_path``: ``calibration/perspective.txt`` (``P_rect_0k`` / ``R_rect_0k``), ``calibration/image_02.yaml``,
``image_03.yaml`` (OpenCV-style yaml of the MEI model whose first line is not yaml), ``calibration/calib_cam_to_pose.txt``
(``image_0k: 12 numbers``), ``data_poses/<seq>/poses.txt`` (``frame 12 numbers`` per key frame, base-link -> world),
``data_2d_raw/<seq>/image_00|image_01/data_rect/%010d.png`` and ``.../image_02|image_03/data_rgb/%010d.png``.
The split ("meta") file has ``sequence,pose_index,image_index,former_image_index,latter_image_index`` per line.
A sample: ('image', f) uint8 frames (+ ('original_image', f) copies for the perspective cameras), 'P2' 3x4 float32 with the
3x3 intrinsics and its copy 'original_P2', ('relative_pose', f) camera-frame motion, an all-ones fp64 'patched_mask' (or the
fisheye validity mask), for the fisheye cameras also 'calib_meta' = the yaml dict (xi, k1, k2 ... read by FishEyeDecoder);
then the configured augmentation.
"""
import os
from copy import deepcopy

import cv2
import numpy as np
import torch.utils.data

from ..utils.builder import build
from .kitti import read_image


def cam_relative_pose_nusc(T_imu2world_0, T_imu2world_1, T_imu2cam):
    """cam <- base_1 <- world <- base_0 <- cam (utils.py:63-64)."""
    return T_imu2cam @ np.linalg.inv(T_imu2world_1) @ T_imu2world_0 @ np.linalg.inv(T_imu2cam)


def read_extrinsic_from_sequence(path):
    """calib_cam_to_pose.txt -> {'T_image0'..'T_image3'} 4x4 camera -> base-link."""
    out = {f"T_image{k}": np.eye(4) for k in range(4)}
    with open(path) as f:
        for line in f:
            for k in range(4):
                if line.startswith(f"image_0{k}"):
                    vals = line.strip().split(" ")
                    out[f"T_image{k}"][:3, :] = np.array([float(x) for x in vals[1:13]]).reshape(3, 4)
    return out


def read_fisheycalib(path):
    import yaml
    with open(path) as f:
        f.readline()                      # "%YAML:1.0": not standard yaml
        return yaml.safe_load(f)


def extract_P_from_fisheye_calib(calib):
    pp = calib["projection_parameters"]
    P = np.zeros([3, 4])
    P[0, 0], P[1, 1], P[0, 2], P[1, 2], P[2, 2] = pp["gamma1"], pp["gamma2"], pp["u0"], pp["v0"], 1
    return P


def read_poses_file(path):
    frames, poses = [], []
    with open(path) as f:
        for line in f:
            vals = line.strip().split(" ")
            if len(vals) < 13:
                continue
            frames.append(int(vals[0]))
            T = np.eye(4)
            T[:3, :] = np.array([float(x) for x in vals[1:13]]).reshape(3, 4)
            poses.append(T)
    return frames, np.array(poses)


def read_cam_to_velo(path):
    """calib_cam_to_velo.txt: one line of 12 numbers -> 4x4 (kitti360_dataset.py:73-83, there ``read_T_from_sequence``)."""
    with open(path) as f:
        vals = f.readline().strip().split(" ")
    T = np.eye(4)
    T[:3, :] = np.array([float(x) for x in vals[:12]]).reshape(3, 4)
    return T


def read_P01_from_sequence(path):
    """perspective.txt -> P_rect_00, P_rect_01 (3x4) and R_rect_00, R_rect_01 embedded in 4x4."""
    P, R = {}, {0: np.eye(4), 1: np.eye(4)}
    with open(path) as f:
        for line in f:
            vals = line.strip().split(" ")
            for k in (0, 1):
                if line.startswith(f"P_rect_0{k}"):
                    P[k] = np.array([float(x) for x in vals[1:13]]).reshape(3, 4)
                if line.startswith(f"R_rect_0{k}"):
                    R[k][:3, :3] = np.array([float(x) for x in vals[1:10]]).reshape(3, 3)
    assert 0 in P, f"can not find P0 in file {path}"
    assert 1 in P, f"can not find P1 in file {path}"
    return P[0], P[1], R[0], R[1]


class _Kitti360Base(torch.utils.data.Dataset):
    """Meta-file parsing, key-frame poses, the static / jump filter and the random camera choice shared by both readers."""
    CAMERAS = None            # (left dir, right dir), image sub-directory

    def __init__(self, **data_cfg):
        super().__init__()
        cfg = data_cfg
        self.raw_path = cfg.get("raw_path", "/data/KITTI-360")
        self.meta_file = cfg.get("split_file", "kitti360_meta.txt")
        self._set_dirs(cfg)
        self.pose_dir = os.path.join(self.raw_path, "data_poses")
        self.pc_dir = os.path.join(self.raw_path, "data_3d_raw")
        self.frame_ids = list(cfg.get("frame_ids", [0, -1, 1]))
        self.imdb, self.sequence_names = [], set()
        with open(self.meta_file) as f:
            for line in f:
                if not line.strip():
                    continue
                seq, pose_index, img_index, former, latter = line.strip().split(",")
                self.sequence_names.add(seq)
                by_frame = {0: int(img_index), -1: int(former), 1: int(latter)}
                self.imdb.append(dict(sequence_name=seq, pose_indexes=[int(pose_index) + i for i in self.frame_ids],
                                      img_indexes=[by_frame[i] for i in self.frame_ids]))
        self._load_calib()
        self.keypose = {seq: read_poses_file(os.path.join(self.pose_dir, seq, "poses.txt"))[1] for seq in self.sequence_names}
        self.is_motion_mask = cfg.get("is_motion_mask", False)
        self.precompute_path = cfg.get("motion_mask_path", "")
        self.is_filter_static = cfg.get("is_filter_static", True)
        self.filter_threshold = cfg.get("filter_threshold", 0.03)
        if self.is_filter_static:
            self.imdb = self._filter_indexes()
        self.use_right_image = cfg.get("use_right_image", True)
        self._post_init(cfg)
        self.transform = build(**cfg["augmentation"])

    def _set_dirs(self, cfg):
        self.img_dir, self.calib_dir = os.path.join(self.raw_path, "data_2d_raw"), os.path.join(self.raw_path, "calibration")

    def _post_init(self, cfg):
        pass

    def _relative(self, poses, k, extrinsics):
        return cam_relative_pose_nusc(poses[0], poses[k], np.linalg.inv(extrinsics)).astype(np.float32)

    def _filter_indexes(self):
        print(f"Start Filtering indexes, original length {len(self)}")
        keep = []
        ext = self.cam_calib["T_rect02baselink"]
        for obj in self.imdb:
            poses = self.keypose[obj["sequence_name"]][obj["pose_indexes"]]
            moves = [np.linalg.norm(self._relative(poses, k + 1, ext)[:3, 3]) for k in range(len(self.frame_ids) - 1)]
            if all(self.filter_threshold <= m <= 3 for m in moves):
                keep.append(obj)
        print(f"Finished filtering indexes, find dynamic instances {len(keep)}")
        return keep

    def __len__(self):
        return len(self.imdb)

    def _choose_camera(self):
        """0 = left, 1 = right (one draw from the global numpy stream when both are allowed)."""
        return 0 if ((not self.use_right_image) or (np.random.rand() < 0.5)) else 1

    def _frames(self, obj, side, with_original):
        ext = self.cam_calib["T_rect02baselink" if side == 0 else "T_rect12baselink"]
        data = {}
        poses = self.keypose[obj["sequence_name"]][obj["pose_indexes"]]
        for k, f in enumerate(self.frame_ids[1:]):
            data[("relative_pose", f)] = self._relative(poses, k + 1, ext)
        image_dir = os.path.join(self.img_dir, obj["sequence_name"], self.CAMERAS[0][side], self.CAMERAS[1])
        for f, i in zip(self.frame_ids, obj["img_indexes"]):
            data[("image", f)] = read_image(os.path.join(image_dir, f"{i:010d}.png"))
            if with_original:
                data[("original_image", f)] = data[("image", f)].copy()
        P = self.cam_calib["P0" if side == 0 else "P1"]
        data["P2"] = np.zeros((3, 4), dtype=np.float32)
        data["P2"][0:3, 0:3] = P[0:3, 0:3]
        data["original_P2"] = data["P2"].copy()
        return data


class KITTI360MonoDataset(_Kitti360Base):
    """Perspective cameras image_00 / image_01 (rectified).  Keywords: raw_path, split_file, frame_ids ([0, -1, 1]),
    is_filter_static / filter_threshold (drops samples whose camera moves less than the threshold or more than 3 m to a
    neighbour), use_right_image (random left / right camera), is_motion_mask, augmentation."""
    CAMERAS = (("image_00", "image_01"), "data_rect")

    def _load_calib(self):
        P0, P1, R0, R1 = read_P01_from_sequence(os.path.join(self.calib_dir, "perspective.txt"))
        ext = read_extrinsic_from_sequence(os.path.join(self.calib_dir, "calib_cam_to_pose.txt"))
        self.cam_calib = dict(P0=P0, P1=P1, T_rect02baselink=R0 @ ext["T_image0"], T_rect12baselink=R1 @ ext["T_image1"])

    def __getitem__(self, index):
        obj = self.imdb[index]
        data = self._frames(obj, self._choose_camera(), with_original=True)
        h, w = data[("image", 0)].shape[:2]
        data["patched_mask"] = np.ones([h, w])
        return self.transform(deepcopy(data))


class KITTI360FisheyeDataset(_Kitti360Base):
    """Fisheye cameras image_02 / image_03.  Extra keywords: resized_root (pre-resized images + calibration),
    fisheye_mask (path of the validity-mask image)."""
    CAMERAS = (("image_02", "image_03"), "data_rgb")

    def _set_dirs(self, cfg):
        self.resized_root = cfg.get("resized_root")
        if self.resized_root is not None:
            self.img_dir, self.calib_dir = self.resized_root, os.path.join(self.resized_root, "calibration")
        else:
            super()._set_dirs(cfg)

    def _post_init(self, cfg):
        mask_path = cfg.get("fisheye_mask")
        # (the reference ignores the configured path and reads a hard-coded one, fisheye_dataset.py:163; the path is honoured here)
        self.fish_eye_mask = cv2.imread(mask_path, -1) if mask_path is not None else None

    def _load_calib(self):
        left = read_fisheycalib(os.path.join(self.calib_dir, "image_02.yaml"))
        right = read_fisheycalib(os.path.join(self.calib_dir, "image_03.yaml"))
        ext = read_extrinsic_from_sequence(os.path.join(self.calib_dir, "calib_cam_to_pose.txt"))
        self.cam_calib = dict(P0=extract_P_from_fisheye_calib(left), P1=extract_P_from_fisheye_calib(right),
                              T_rect02baselink=ext["T_image2"], T_rect12baselink=ext["T_image3"], left_meta=left, right_meta=right)

    def __getitem__(self, index):
        obj = self.imdb[index]
        side = self._choose_camera()
        data = self._frames(obj, side, with_original=False)
        data["calib_meta"] = deepcopy(self.cam_calib["left_meta" if side == 0 else "right_meta"])
        h, w = data[("image", 0)].shape[:2]
        if self.fish_eye_mask is not None:
            data["patched_mask"] = cv2.resize(self.fish_eye_mask, (w, h), interpolation=cv2.INTER_NEAREST)
        else:
            data["patched_mask"] = np.ones([h, w])
        return self.transform(deepcopy(data))
