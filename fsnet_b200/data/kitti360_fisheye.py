"""KITTI-360 fisheye reader with the reference's class name, keywords and sample schema
(monodepth/data/datasets/fisheye_dataset.py:17-262): the caller of the MEI-camera loss head (FishEyeDecoder).

Layout under ``raw_path``: ``calibration/image_02.yaml``, ``image_03.yaml`` (OpenCV-style yaml of the MEI model whose
first line is not yaml), ``calibration/calib_cam_to_pose.txt`` (``image_0k: 12 numbers``), ``data_poses/<seq>/poses.txt``
(``frame 12 numbers`` per key frame, base-link -> world), ``data_2d_raw/<seq>/image_02|image_03/data_rgb/%010d.png``.
The split ("meta") file has ``sequence,pose_index,image_index,former_image_index,latter_image_index`` per line.
A sample: ('image', f) uint8 frames, 'P2' = [[gamma1,0,u0,0],[0,gamma2,v0,0],[0,0,1,0]] float32 and its copy
'original_P2', 'calib_meta' = the yaml dict (xi, k1, k2 ... read by the head), ('relative_pose', f) camera-frame motion,
'patched_mask' (the fisheye validity mask resized with nearest neighbour, or ones); then the configured augmentation.
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


class KITTI360FisheyeDataset(torch.utils.data.Dataset):
    """Keywords: raw_path, split_file, resized_root, frame_ids ([0, -1, 1]), is_filter_static / filter_threshold (drops
    samples whose camera moves less than the threshold or more than 3 m to a neighbour), use_right_image (random left /
    right camera, drawn from the global numpy stream), fisheye_mask (path of the validity mask image), is_motion_mask,
    augmentation."""

    def __init__(self, **data_cfg):
        super().__init__()
        cfg = data_cfg
        self.raw_path = cfg.get("raw_path", "/data/KITTI-360")
        self.meta_file = cfg.get("split_file", "kitti360_meta.txt")
        self.resized_root = cfg.get("resized_root")
        if self.resized_root is not None:
            self.img_dir, self.calib_dir = self.resized_root, os.path.join(self.resized_root, "calibration")
        else:
            self.img_dir, self.calib_dir = os.path.join(self.raw_path, "data_2d_raw"), os.path.join(self.raw_path, "calibration")
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
        mask_path = cfg.get("fisheye_mask")
        # (the reference ignores the configured path and reads a hard-coded one, fisheye_dataset.py:163; the path is honoured here)
        self.fish_eye_mask = cv2.imread(mask_path, -1) if mask_path is not None else None
        self.transform = build(**cfg["augmentation"])

    def _load_calib(self):
        left = read_fisheycalib(os.path.join(self.calib_dir, "image_02.yaml"))
        right = read_fisheycalib(os.path.join(self.calib_dir, "image_03.yaml"))
        ext = read_extrinsic_from_sequence(os.path.join(self.calib_dir, "calib_cam_to_pose.txt"))
        self.cam_calib = dict(P0=extract_P_from_fisheye_calib(left), P1=extract_P_from_fisheye_calib(right),
                              T_rect02baselink=ext["T_image2"], T_rect12baselink=ext["T_image3"], left_meta=left, right_meta=right)

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

    def __getitem__(self, index):
        obj = self.imdb[index]
        seq = obj["sequence_name"]
        if (not self.use_right_image) or (np.random.rand() < 0.5):
            ext, cam_dir, P, meta = self.cam_calib["T_rect02baselink"], "image_02", self.cam_calib["P0"], self.cam_calib["left_meta"]
        else:
            ext, cam_dir, P, meta = self.cam_calib["T_rect12baselink"], "image_03", self.cam_calib["P1"], self.cam_calib["right_meta"]
        data = {}
        poses = self.keypose[seq][obj["pose_indexes"]]
        for k, f in enumerate(self.frame_ids[1:]):
            data[("relative_pose", f)] = self._relative(poses, k + 1, ext)
        image_dir = os.path.join(self.img_dir, seq, cam_dir, "data_rgb")
        for f, i in zip(self.frame_ids, obj["img_indexes"]):
            data[("image", f)] = read_image(os.path.join(image_dir, f"{i:010d}.png"))
        data["P2"] = np.zeros((3, 4), dtype=np.float32)
        data["P2"][0:3, 0:3] = P[0:3, 0:3]
        data["original_P2"] = data["P2"].copy()
        data["calib_meta"] = deepcopy(meta)
        h, w = data[("image", 0)].shape[:2]
        if self.fish_eye_mask is not None:
            data["patched_mask"] = cv2.resize(self.fish_eye_mask, (w, h), interpolation=cv2.INTER_NEAREST)
        else:
            data["patched_mask"] = np.ones([h, w])
        return self.transform(deepcopy(data))
