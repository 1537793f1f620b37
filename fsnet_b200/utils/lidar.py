"""LiDAR ground truth for the depth evaluators: KITTI calibration text files, Velodyne .bin scans and their z-buffered
projection into a camera (monodepth/networks/utils/monodepth_utils.py:291-459).

Semantics kept from the reference: points with forward coordinate < 0 are dropped before projection; pixel coordinates
are ``round(u) - 1`` (the KITTI MATLAB convention); when several points fall on one pixel the NEAREST wins; negative
depths become 0 ("no measurement").  The z-buffer is one ``np.minimum.at`` scatter instead of a Python loop over duplicates.
"""
import os

import numpy as np

_FLOAT_CHARS = set("0123456789.e+- ")


def read_calib_file(path):
    """``key: v0 v1 ...`` lines -> {key: float array}; non-numeric values stay strings (monodepth_utils.py:339-358)."""
    out = {}
    with open(path) as f:
        for line in f:
            if ":" not in line:
                continue
            key, value = line.split(":", 1)
            value = value.strip()
            out[key] = value
            if _FLOAT_CHARS.issuperset(value):
                try:
                    out[key] = np.array([float(v) for v in value.split(" ")])
                except ValueError:
                    pass
    return out


def load_velodyne_points(filename):
    """[N,4] float32 (forward, left, up, 1) (monodepth_utils.py:360-366)."""
    pts = np.fromfile(filename, dtype=np.float32).reshape(-1, 4)
    pts[:, 3] = 1.0
    return pts


def sub2ind(shape, row, col):
    """MATLAB-style linear index used by the reference to find duplicates (monodepth_utils.py:291-295)."""
    m, n = shape
    return row * (n - 1) + col - 1


def _zbuffer(points_h, depth_of, P_velo2im, im_shape):
    """points_h [N,4] homogeneous Velodyne points (already filtered to forward >= 0), ``depth_of`` the value stored per point
    (None: the projective z).  Returns the [h, w] float64 depth image."""
    h, w = int(im_shape[0]), int(im_shape[1])
    uvz = points_h @ np.asarray(P_velo2im).T
    z = uvz[:, 2] if depth_of is None else depth_of
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.round(uvz[:, 0] / uvz[:, 2]) - 1
        v = np.round(uvz[:, 1] / uvz[:, 2]) - 1
    ok = (u >= 0) & (v >= 0) & (u < w) & (v < h)
    u, v, z = u[ok].astype(np.int64), v[ok].astype(np.int64), z[ok]
    depth = np.full((h, w), np.inf)
    np.minimum.at(depth, (v, u), z)
    depth[~np.isfinite(depth)] = 0
    depth[depth < 0] = 0
    return depth


def generate_depth_map(calib_dir, velo_filename, cam=2, vel_depth=False):
    """KITTI raw: Velodyne scan -> depth image of rectified camera ``cam`` (monodepth_utils.py:368-420).
    ``vel_depth``: store the Velodyne forward coordinate instead of the camera z (what the Eigen protocol evaluates)."""
    cam2cam = read_calib_file(os.path.join(calib_dir, "calib_cam_to_cam.txt"))
    velo2cam = read_calib_file(os.path.join(calib_dir, "calib_velo_to_cam.txt"))
    T = np.eye(4)
    T[:3, :3] = velo2cam["R"].reshape(3, 3)
    T[:3, 3] = velo2cam["T"]
    im_shape = cam2cam["S_rect_02"][::-1].astype(np.int32)
    R = np.eye(4)
    R[:3, :3] = cam2cam["R_rect_00"].reshape(3, 3)
    P_velo2im = cam2cam["P_rect_0" + str(cam)].reshape(3, 4) @ R @ T
    velo = load_velodyne_points(velo_filename)
    velo = velo[velo[:, 0] >= 0]
    return _zbuffer(velo, velo[:, 0] if vel_depth else None, P_velo2im, im_shape)


def project_depth_map(velo, P_velo2im, im_shape):
    """Same projection with the matrix given (KITTI-360; always Velodyne-forward depth) (monodepth_utils.py:422-459)."""
    pts = np.array(velo[velo[:, 0] >= 0], copy=True)
    pts[:, 3] = 1.0
    return _zbuffer(pts, pts[:, 0], P_velo2im, im_shape)
