"""nuScenes reader over a pre-exported JSON index, with the reference's class name, keywords and sample schema
(monodepth/data/datasets/nuscene_dataset.py:167-238, ``NusceneJsonDataset`` -- the one the nuScenes configs build).

The JSON holds ``{"samples": [{frame0, frame1, frame-1: image paths, P2: 9 numbers, pose01 / pose0-1: 16 numbers,
camera_type, camera_type_indexes}, ...]}``.  A sample: ('image', f) / ('original_image', f) uint8 frames, 'P2' 3x4 float32
and 'original_P2', ('relative_pose', +-1), an fp64 'patched_mask' of ones (rows 700.. zeroed for CAM_BACK: the ego car),
'camera_type', 'camera_type_index', ('filename', 0) = the last three path components of the target frame, optionally
('vo_depth', 0); then the configured augmentation."""
import json
import os
from copy import deepcopy

import cv2
import numpy as np
import torch.utils.data

from ..utils.builder import build
from .kitti import read_image


def read_vo_depth(image_path):
    """16-bit visual-odometry depth png -> metres; anything outside 3..80 m is pushed to 120 (utils.py:12-19)."""
    depth = cv2.imread(image_path, -1) / 65535.0 * 120
    depth[depth < 3] = 120
    depth[depth > 80] = 120
    return depth


class NusceneJsonDataset(torch.utils.data.Dataset):
    CAMERAS = ["CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_BACK_RIGHT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_FRONT_LEFT"]

    def __init__(self, **data_cfg):
        super().__init__()
        cfg = data_cfg
        self.json_path = cfg.get("json_path", "/home/monodepth/meta_data/nusc_trainsub/json_nusc_front_train.json")
        with open(self.json_path) as f:
            self.json_dict = json.load(f)
        self.image_keys = list(cfg.get("image_keys", ["frame0", "frame1", "frame-1"]))
        self.pose_keys = list(cfg.get("pose_keys", ["pose01", "pose0-1"]))
        self.intrinsic_key = cfg.get("intrinsic_key", "P2")
        self.cameras = list(cfg.get("channels", self.CAMERAS))
        self.frame_ids = list(cfg.get("frame_ids", [0, 1, -1]))
        self.transform = build(**cfg["augmentation"])
        self.vo_path = cfg.get("vo_path")
        self.is_read_vo_depth = self.vo_path is not None

    def __len__(self):
        return len(self.json_dict["samples"])

    def __getitem__(self, index):
        sample = self.json_dict["samples"][index]
        frames = [read_image(sample[key]) for key in self.image_keys]
        data = {("relative_pose", 1): np.array(sample["pose01"]).reshape(4, 4).astype(np.float32),
                ("relative_pose", -1): np.array(sample["pose0-1"]).reshape(4, 4).astype(np.float32)}
        for frame, f in zip(frames, self.frame_ids):
            data[("image", f)] = frame
            data[("original_image", f)] = frame.copy()
        h, w = data[("image", 0)].shape[:2]
        data["patched_mask"] = np.ones([h, w])
        if sample["camera_type"] == "CAM_BACK":
            data["patched_mask"][700:, :] = 0
        data["P2"] = np.zeros((3, 4), dtype=np.float32)
        data["P2"][0:3, 0:3] = np.array(sample[self.intrinsic_key]).reshape(3, 3).astype(np.float32)
        data["original_P2"] = data["P2"].copy()
        data["camera_type_index"] = sample["camera_type_indexes"]
        data[("filename", 0)] = os.path.join(*sample[self.image_keys[0]].split("/")[-3:])
        data["camera_type"] = sample["camera_type"]
        if self.is_read_vo_depth:
            vo_file = data[("filename", 0)].replace("samples", self.vo_path).replace(".jpg", ".png")
            if os.path.isfile(vo_file):
                data[("vo_depth", 0)] = read_vo_depth(vo_file)
            else:
                print(f"No VO Depth file found at {index}, {vo_file}")
        return self.transform(deepcopy(data))
