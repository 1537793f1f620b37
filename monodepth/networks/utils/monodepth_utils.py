"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.networks.pose_decoder import (rot_from_axisangle, get_translation_matrix,  # noqa: F401
                                              transformation_from_parameters)
from fsnet_b200.utils.metrics import (compute_depth_errors, compute_errors, depth_to_disp, disp_to_depth,  # noqa: F401
                                      inverse_sigmoid)
from fsnet_b200.utils.lidar import (generate_depth_map, load_velodyne_points, project_depth_map, read_calib_file,  # noqa: F401
                                    sub2ind)
