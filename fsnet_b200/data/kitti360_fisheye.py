"""Kept for import compatibility: the KITTI-360 readers live in fsnet_b200/data/kitti360.py."""
from .kitti360 import *  # noqa: F401,F403
from .kitti360 import KITTI360FisheyeDataset, cam_relative_pose_nusc, extract_P_from_fisheye_calib, read_extrinsic_from_sequence, read_fisheycalib, read_poses_file  # noqa: F401
