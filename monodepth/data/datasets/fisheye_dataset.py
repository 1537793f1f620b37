"""Reference-compatible dotted names; the implementation lives in fsnet_b200."""
from fsnet_b200.data.kitti360_fisheye import (KITTI360FisheyeDataset, extract_P_from_fisheye_calib, read_extrinsic_from_sequence,  # noqa: F401
                                              read_fisheycalib, read_poses_file)
