"""Eigen-protocol depth evaluators with the reference's names and keywords (monodepth/evaluation/kitti_unsupervised_eval.py).

``KittiEigenEvaluator(data_path, split_file, gt_saved_file, is_evaluate_absolute=False)``: loads the LiDAR ground truth
from ``gt_saved_file`` (npz, key "data") or projects it from the raw Velodyne scans listed in ``split_file`` and saves it;
``single_call(depth, index)`` -> {ratio, error[7], abs_error[7]} (Garg/Eigen crop, 1e-3..80 m, median scaling);
``log(writer, mean_errors, mean_abs_errors, ...)`` prints the reference's table; ``__call__(result_path, ...)`` evaluates a
directory of 16-bit depth pngs.  ``Kitti360Evaluator`` differs only in how the ground truth is projected."""
import os

import cv2
import numpy as np
from PIL import Image

from ..data.kitti import read_depth
from ..data.kitti360 import read_P01_from_sequence, read_cam_to_velo
from ..utils.lidar import generate_depth_map, project_depth_map
from ..utils.metrics import compute_errors

_NAMES = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")
MIN_DEPTH, MAX_DEPTH = 1e-3, 80.0


def eigen_crop_mask(gt_depth):
    """valid LiDAR returns (1e-3 < d < 80) inside the Garg/Eigen crop (kitti_unsupervised_eval.py:50-56)."""
    h, w = gt_depth.shape[:2]
    y0, y1, x0, x1 = np.array([0.40810811 * h, 0.99189189 * h, 0.03594771 * w, 0.96405229 * w]).astype(np.int32)
    mask = np.zeros(gt_depth.shape, dtype=bool)
    mask[y0:y1, x0:x1] = True
    return mask & (gt_depth > MIN_DEPTH) & (gt_depth < MAX_DEPTH)


def _table(title, values):
    return (title + "\n  " + ("{:>8} | " * 7).format(*_NAMES) + "\n" + ("&{: 8.3f}  " * 7).format(*np.asarray(values).tolist()) + "\\\\")


class KittiEigenEvaluator(object):
    def __init__(self, data_path, split_file, gt_saved_file, is_evaluate_absolute=False):
        self.is_evaluate_absolute = is_evaluate_absolute
        if os.path.isfile(gt_saved_file):
            self.gt_depths = np.load(gt_saved_file, fix_imports=True, encoding="latin1", allow_pickle=True)["data"]
        else:
            print(f"Start exporting ground truth depths specified by {split_file} to {gt_saved_file}")
            self._precompute(data_path, split_file, gt_saved_file)

    @staticmethod
    def _save(gt_saved_file, gt_depths):
        same = len({g.shape for g in gt_depths}) <= 1
        data = np.array(gt_depths) if same else np.array(gt_depths + [None], dtype=object)[:-1]   # ragged image sizes
        np.savez_compressed(gt_saved_file, data=data)

    def _precompute(self, data_path, split_file, gt_saved_file):
        gt_depths = []
        with open(split_file) as f:
            for line in f:
                if not line.strip():
                    continue
                folder, frame_id = line.split()[:2]
                velo = os.path.join(data_path, folder, "velodyne_points/data", "{:010d}.bin".format(int(frame_id)))
                gt_depths.append(generate_depth_map(os.path.join(data_path, folder.split("/")[0]), velo, 2, True).astype(np.float32))
        self._save(gt_saved_file, gt_depths)
        self.gt_depths = gt_depths

    def _single_loss(self, depth_0, gt_depth):
        gt_depth = np.asarray(gt_depth)
        h, w = gt_depth.shape[:2]
        mask = eigen_crop_mask(gt_depth)
        pred = cv2.resize(depth_0, (w, h))[mask]
        gt = gt_depth[mask]
        if len(pred) == 0:
            raise ValueError
        ratio = np.median(gt) / np.median(pred)
        error = compute_errors(gt, np.clip(pred * ratio, MIN_DEPTH, MAX_DEPTH))
        abs_error = compute_errors(gt, np.clip(pred, MIN_DEPTH, MAX_DEPTH))
        return dict(ratio=ratio, error=error, abs_error=abs_error)

    def single_call(self, depth_0, index):
        return self._single_loss(depth_0, self.gt_depths[index])

    def log(self, writer, mean_errors, mean_abs_errors, global_step=0, epoch_num=0, is_print=True):
        log_str = _table(f"Epoch {epoch_num}", mean_errors) + "\n" + _table(f"Epoch {epoch_num}| Abs Error without Scaled", mean_abs_errors)
        if writer is not None:
            writer.add_text("evaluation logs", log_str.replace(" ", "&nbsp;").replace("\n", "  \n"), global_step=epoch_num)
        if is_print:
            print(log_str)
        return log_str

    def __call__(self, result_path, writer=None, global_step=0, epoch_num=0):
        files = sorted(os.listdir(result_path))
        if len(files) != len(self.gt_depths):
            print(f"The length of pred_depths is {len(files)} while the length of gt_depths is {len(self.gt_depths)}")
            print("Drop evaluation")
            return None
        res = [self._single_loss(read_depth(os.path.join(result_path, f)), self.gt_depths[i]) for i, f in enumerate(files)]
        scales = np.array([r["ratio"] for r in res])
        mean_errors = np.array([r["error"] for r in res]).mean(0)
        mean_abs_errors = np.array([r["abs_error"] for r in res]).mean(0)
        log_str = (_table(f"Epoch {epoch_num} | Scaled Error | {scales.mean()}, {scales.std()}", mean_errors) + "\n"
                   + _table(f"Epoch {epoch_num} | Abs Error without Scaled", mean_abs_errors))
        if writer is not None:
            writer.add_text("evaluation logs", log_str.replace(" ", "&nbsp;").replace("\n", "  \n"), global_step=epoch_num)
        print(log_str)
        return dict(error=mean_errors, abs_error=mean_abs_errors, ratios=scales)


class Kitti360Evaluator(KittiEigenEvaluator):
    """KITTI-360: split lines ``sequence,pose_index,image_index,former,latter``; the scan of ``image_index`` is projected
    with P_rect_00 @ R_rect_00 @ inv(T_cam0->velo) into the size of the rectified image (kitti_unsupervised_eval.py:163-212)."""

    def _load_calib(self, calib_dir):
        P0, _, R0, _ = read_P01_from_sequence(os.path.join(calib_dir, "perspective.txt"))
        self.cam_calib = dict(P0=P0, R0=R0, T_cam2velo=read_cam_to_velo(os.path.join(calib_dir, "calib_cam_to_velo.txt")))

    def _precompute(self, data_path, split_file, gt_saved_file):
        self._load_calib(os.path.join(data_path, "calibration"))
        P_velo2img = self.cam_calib["P0"] @ self.cam_calib["R0"] @ np.linalg.inv(self.cam_calib["T_cam2velo"])
        gt_depths = []
        with open(split_file) as f:
            for line in f:
                if not line.strip():
                    continue
                sequence, _, img_index = line.strip().split(",")[:3]
                name = "{:010d}".format(int(img_index))
                velo = np.fromfile(os.path.join(data_path, "data_3d_raw", sequence, "velodyne_points/data", name + ".bin"),
                                   dtype=np.float32).reshape(-1, 4)
                with Image.open(os.path.join(data_path, "data_2d_raw", sequence, "image_00", "data_rect", name + ".png")) as im:
                    shape = np.array(im.size)[::-1].astype(np.int32)
                gt_depths.append(project_depth_map(velo, P_velo2img, shape).astype(np.float32))
        self._save(gt_saved_file, gt_depths)
        self.gt_depths = gt_depths
