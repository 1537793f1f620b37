"""Reference-compatible dotted names; the implementation lives in fsnet_b200."""
from fsnet_b200.data import kitti360 as _k
from fsnet_b200.data.kitti360 import KITTI360MonoDataset, read_P01_from_sequence, read_poses_file  # noqa: F401


def read_extrinsic_from_sequence(file):
    """(T_image0, T_image1) like kitti360_dataset.py:42-58 (the fisheye module's variant returns all four as a dict)."""
    ext = _k.read_extrinsic_from_sequence(file)
    return ext["T_image0"], ext["T_image1"]
