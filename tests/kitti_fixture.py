"""A miniature KITTI-raw directory tree (two drives, both cameras, calibration files, devkit poses, an Eigen-style split
file, 16-bit depth maps) written deterministically into a temporary directory: lets the dataset readers be exercised --
by the reference (golden generation) and by this repo (tests) -- without KITTI on disk."""
import os

import cv2
import numpy as np
import scipy.io as sio
from PIL import Image

DATE = "2011_09_26"
DRIVES = ["2011_09_26_drive_0001_sync", "2011_09_26_drive_0002_sync"]
H, W, N = 120, 400, 8


def _rigid(g, scale):
    ang = g.uniform(-0.02, 0.02, size=3)
    cx, cy, cz = np.cos(ang)
    sx, sy, sz = np.sin(ang)
    R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    return R, g.uniform(-scale, scale, size=3)


def build_tree(root, seed=5):
    g = np.random.default_rng(seed)
    d = os.path.join(root, "raw", DATE)
    os.makedirs(d, exist_ok=True)
    P2 = np.array([[0.58 * W, 0, 0.5 * W, 4.4], [0, 1.92 * H, 0.5 * H, 0.2], [0, 0, 1, 0.003]])
    P3 = P2.copy(); P3[0, 3] = -33.0
    with open(os.path.join(d, "calib_cam_to_cam.txt"), "w") as f:
        f.write("calib_time: 09-Jan-2012 13:57:47\n")
        f.write("P_rect_02: " + " ".join(f"{v:.9e}" for v in P2.reshape(-1)) + "\n")
        f.write("P_rect_03: " + " ".join(f"{v:.9e}" for v in P3.reshape(-1)) + "\n")
    for name, rt in (("calib_velo_to_cam.txt", _rigid(g, 0.3)), ("calib_imu_to_velo.txt", _rigid(g, 0.8))):
        with open(os.path.join(d, name), "w") as f:
            f.write("calib_time: 25-May-2012 16:47:16\n")
            f.write("R: " + " ".join(f"{v:.9e}" for v in rt[0].reshape(-1)) + "\n")
            f.write("T: " + " ".join(f"{v:.9e}" for v in rt[1]) + "\n")
    lines = []
    for di, drive in enumerate(DRIVES):
        for cam in ("image_02", "image_03"):
            os.makedirs(os.path.join(d, drive, cam, "data"), exist_ok=True)
        os.makedirs(os.path.join(d, drive, "oxts"), exist_ok=True)
        os.makedirs(os.path.join(d, drive, "depth"), exist_ok=True)
        poses = np.tile(np.eye(4), (N, 1, 1))
        for k in range(1, N):
            R, t = _rigid(g, 0.0)
            step = np.eye(4)
            step[:3, :3] = R
            # drive 0002 stands still between frames 3 and 4 (the static filter must drop the samples around them)
            step[:3, 3] = [0.0 if (di == 1 and k == 4) else 0.9, 0.01, 0.0]
            poses[k] = poses[k - 1] @ step
        sio.savemat(os.path.join(d, drive, "oxts", "pose.mat"), {"pose_mat": poses})
        for k in range(N):
            for cam in ("image_02", "image_03"):
                lo = g.integers(0, 256, size=(H // 8, W // 8, 3)).astype(np.uint8)
                img = np.kron(lo, np.ones((8, 8, 1), dtype=np.uint8)) // 2 + g.integers(0, 128, size=(H, W, 3)).astype(np.uint8)
                Image.fromarray(img).save(os.path.join(d, drive, cam, "data", "%010d.png" % k))
            depth = (g.uniform(0, 80, size=(H, W)) * (g.uniform(size=(H, W)) < 0.1) * 256).astype(np.uint16)
            cv2.imwrite(os.path.join(d, drive, "depth", "%010d.png" % k), depth)
        for k in range(1, N - 1):
            lines.append(f"{DATE}/{drive} {k} {'l' if (k + di) % 2 == 0 else 'r'}")
    split = os.path.join(root, "split.txt")
    with open(split, "w") as f:
        f.write("\n".join(lines) + "\n")
    return os.path.join(root, "raw"), split


# --------------------------------------------------------------------------------------------------
# miniature KITTI-360 (fisheye cameras)
# --------------------------------------------------------------------------------------------------
SEQ360 = "2013_05_28_drive_0000_sync"
FH = FW = 96


def build_kitti360_tree(root, seed=9):
    g = np.random.default_rng(seed)
    raw = os.path.join(root, "kitti360")
    calib = os.path.join(raw, "calibration")
    os.makedirs(calib, exist_ok=True)
    for cam, u0 in (("image_02", 0.512), ("image_03", 0.498)):
        with open(os.path.join(calib, cam + ".yaml"), "w") as f:
            f.write("%YAML:1.0\n---\nmodel_type: MEI\ncamera_name: " + cam + f"\nimage_width: {FW}\nimage_height: {FH}\n"
                    "mirror_parameters:\n   xi: 2.2134047507854890e+00\n"
                    "distortion_parameters:\n   k1: 1.6798235660113681e-02\n   k2: 1.6548773243373522e+00\n   p1: 4.2e-04\n   p2: 4.2e-04\n"
                    f"projection_parameters:\n   gamma1: {1336.3 * FW / 1400:.6f}\n   gamma2: {1335.8 * FH / 1400:.6f}\n"
                    f"   u0: {u0 * FW:.6f}\n   v0: {0.504 * FH:.6f}\n")
    with open(os.path.join(calib, "calib_cam_to_pose.txt"), "w") as f:
        for k in range(4):
            R, t = _rigid(g, 0.6)
            T = np.concatenate([R, t[:, None]], 1)
            f.write(f"image_0{k}: " + " ".join(f"{v:.9e}" for v in T.reshape(-1)) + "\n")
    os.makedirs(os.path.join(raw, "data_poses", SEQ360), exist_ok=True)
    pose = np.eye(4)
    with open(os.path.join(raw, "data_poses", SEQ360, "poses.txt"), "w") as f:
        for k in range(10):
            R, _ = _rigid(g, 0.0)
            step = np.eye(4)
            step[:3, :3] = R
            step[:3, 3] = [0.0 if k == 5 else (5.0 if k == 8 else 0.8), 0.02, 0.0]       # one standing pair, one 5 m jump
            pose = pose @ step
            f.write(f"{k * 3} " + " ".join(f"{v:.9e}" for v in pose[:3].reshape(-1)) + "\n")
    for cam in ("image_02", "image_03"):
        d = os.path.join(raw, "data_2d_raw", SEQ360, cam, "data_rgb")
        os.makedirs(d, exist_ok=True)
        for k in range(0, 30, 3):
            lo = g.integers(0, 256, size=(FH // 8, FW // 8, 3)).astype(np.uint8)
            img = np.kron(lo, np.ones((8, 8, 1), dtype=np.uint8)) // 2 + g.integers(0, 128, size=(FH, FW, 3)).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(d, "%010d.png" % k))
    # perspective cameras: rectified intrinsics / rotations and rectified frames
    PH, PW = 64, 208
    with open(os.path.join(calib, "perspective.txt"), "w") as f:
        for k in (0, 1):
            P = np.array([[0.78 * PW, 0, 0.49 * PW, -50.0 * k], [0, 2.1 * PH, 0.51 * PH, 0], [0, 0, 1, 0]])
            R, _ = _rigid(g, 0.0)
            f.write(f"P_rect_0{k}: " + " ".join(f"{v:.9e}" for v in P.reshape(-1)) + "\n")
            f.write(f"R_rect_0{k}: " + " ".join(f"{v:.9e}" for v in R.reshape(-1)) + "\n")
    for cam in ("image_00", "image_01"):
        d = os.path.join(raw, "data_2d_raw", SEQ360, cam, "data_rect")
        os.makedirs(d, exist_ok=True)
        for k in range(0, 30, 3):
            lo = g.integers(0, 256, size=(PH // 8, PW // 8, 3)).astype(np.uint8)
            img = np.kron(lo, np.ones((8, 8, 1), dtype=np.uint8)) // 2 + g.integers(0, 128, size=(PH, PW, 3)).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(d, "%010d.png" % k))
    meta = os.path.join(root, "kitti360_meta.txt")
    with open(meta, "w") as f:
        for p in range(1, 9):
            f.write(f"{SEQ360},{p},{p * 3},{p * 3 - 3},{p * 3 + 3}\n")
    yy, xx = np.mgrid[0:FH, 0:FW]
    mask = (((yy - FH / 2) ** 2 + (xx - FW / 2) ** 2) < (0.48 * FW) ** 2).astype(np.uint8)
    mask_path = os.path.join(root, "fisheye_mask.png")
    cv2.imwrite(mask_path, mask)
    return raw, meta, mask_path


# --------------------------------------------------------------------------------------------------
# miniature nuScenes JSON export
# --------------------------------------------------------------------------------------------------
def build_nusc_json(root, seed=17, n=5, h0=96, w=160):
    import json
    g = np.random.default_rng(seed)
    cams = ["CAM_FRONT", "CAM_BACK", "CAM_FRONT_LEFT"]
    samples = []
    for i in range(n):
        cam = cams[i % len(cams)]
        d = os.path.join(root, "nusc", "samples", cam)
        os.makedirs(d, exist_ok=True)
        paths = {}
        h = 720 if cam == "CAM_BACK" else h0           # tall enough for the ego-car rows (700..) of the rear camera
        for key in ("frame0", "frame1", "frame-1"):
            lo = g.integers(0, 256, size=(h // 8, w // 8, 3)).astype(np.uint8)
            img = np.kron(lo, np.ones((8, 8, 1), dtype=np.uint8)) // 2 + g.integers(0, 128, size=(h, w, 3)).astype(np.uint8)
            p = os.path.join(d, f"n{i:03d}_{key.replace('-', 'm')}.png")
            Image.fromarray(img).save(p)
            paths[key] = p
        poses = {}
        for key, sign in (("pose01", -1.0), ("pose0-1", 1.0)):
            R, _ = _rigid(g, 0.0)
            T = np.eye(4)
            T[:3, :3] = R
            T[:3, 3] = [0.02, 0.01, sign * 0.6]
            poses[key] = [float(v) for v in T.reshape(-1)]
        K = [0.79 * w, 0.0, 0.5 * w, 0.0, 0.79 * w, 0.5 * h, 0.0, 0.0, 1.0]
        samples.append(dict(camera_type=cam, camera_type_indexes=cams.index(cam), P2=K, **paths, **poses))
    path = os.path.join(root, "nusc.json")
    with open(path, "w") as f:
        json.dump(dict(samples=samples), f)
    return path


# --------------------------------------------------------------------------------------------------
# LiDAR scans for the evaluators (separate generators: the trees above stay byte-identical)
# --------------------------------------------------------------------------------------------------
def _scan(g, n=6000):
    """[n,4] float32 (forward, left, up, reflectance); the fixture's velodyne->camera rotations are near identity, so the
    camera looks along 'up'.  Some points lie behind the sensor (forward < 0) and some behind the image plane (up < 0)."""
    pts = np.stack([g.uniform(-5, 40, n), g.uniform(-15, 15, n), g.uniform(-2, 50, n), g.uniform(0, 1, n)], 1)
    return pts.astype(np.float32)


def add_kitti_lidar(raw, seed=31):
    """Adds S_rect_02 / R_rect_00 to calib_cam_to_cam.txt and velodyne_points/data/*.bin to every drive of build_tree()."""
    g = np.random.default_rng(seed)
    d = os.path.join(raw, DATE)
    R, _ = _rigid(g, 0.0)
    with open(os.path.join(d, "calib_cam_to_cam.txt"), "a") as f:
        f.write(f"S_rect_02: {W:.12e} {H:.12e}\n")
        f.write("R_rect_00: " + " ".join(f"{v:.9e}" for v in R.reshape(-1)) + "\n")
    for drive in DRIVES:
        v = os.path.join(d, drive, "velodyne_points", "data")
        os.makedirs(v, exist_ok=True)
        for k in range(N):
            _scan(g).tofile(os.path.join(v, "%010d.bin" % k))


def add_kitti360_lidar(raw, seed=32):
    """Adds calibration/calib_cam_to_velo.txt and data_3d_raw/<seq>/velodyne_points/data/*.bin to build_kitti360_tree()."""
    g = np.random.default_rng(seed)
    R, t = _rigid(g, 0.4)
    with open(os.path.join(raw, "calibration", "calib_cam_to_velo.txt"), "w") as f:
        f.write(" ".join(f"{v:.9e}" for v in np.concatenate([R, t[:, None]], 1).reshape(-1)) + "\n")
    v = os.path.join(raw, "data_3d_raw", SEQ360, "velodyne_points", "data")
    os.makedirs(v, exist_ok=True)
    for k in range(0, 30, 3):
        _scan(g, 4000).tofile(os.path.join(v, "%010d.bin" % k))
