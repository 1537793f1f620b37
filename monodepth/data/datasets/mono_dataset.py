"""Reference-compatible dotted names; the implementation lives in fsnet_b200."""
from fsnet_b200.data.kitti import (KittiDepthMonoDataset, KittiDepthMonoEigenTestDataset, read_P23_from_sequence,  # noqa: F401
                                   read_T_from_sequence, read_imu2velo, read_split_file)
