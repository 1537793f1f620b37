"""Reference-compatible dotted names; the implementation lives in fsnet_b200."""
from fsnet_b200.data.kitti import cam_relative_pose, read_depth, read_image, read_pose_mat  # noqa: F401
from fsnet_b200.data.kitti360_fisheye import cam_relative_pose_nusc  # noqa: F401,E402
