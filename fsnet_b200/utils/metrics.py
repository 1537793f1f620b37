"""Depth <-> disparity conversions and the standard monocular-depth error metrics, host side
(reference: monodepth/networks/utils/monodepth_utils.py:8-29,251-289; the Eigen-crop / median-scaling protocol of
monodepth/evaluation/kitti_unsupervised_eval.py:47-80).  numpy for the evaluators, torch for in-training logging --
both work on whatever device / array their inputs live on; nothing here is on the training hot path."""
import numpy as np
import torch


def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid output in [0, 1] -> (scaled disparity, depth) with depth in [min_depth, max_depth]."""
    min_disp, max_disp = 1 / max_depth, 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp
    return scaled, 1 / scaled


def depth_to_disp(depth, min_depth, max_depth):
    """Inverse of the depth half of ``disp_to_depth``."""
    return (1 / depth - 1 / max_depth) / (1 / min_depth - 1 / max_depth)


def inverse_sigmoid(x):
    return torch.log(x / (1 - x + 1e-8))


def _errors(gt, pred, xp):
    ratio = xp.maximum(gt / pred, pred / gt)
    a = [(ratio < 1.25 ** k) for k in (1, 2, 3)]
    a1, a2, a3 = [(t.float() if xp is torch else t).mean() for t in a]
    diff = gt - pred
    rmse = xp.sqrt((diff ** 2).mean())
    rmse_log = xp.sqrt(((xp.log(gt) - xp.log(pred)) ** 2).mean())
    abs_rel = (xp.abs(diff) / gt).mean()
    sq_rel = (diff ** 2 / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def compute_errors(gt, pred):
    """(abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3) on numpy arrays of valid pixels."""
    return _errors(np.asarray(gt), np.asarray(pred), np)


def compute_depth_errors(gt, pred):
    """The same seven metrics on torch tensors."""
    return _errors(gt, pred, torch)


def eigen_median_scaled_errors(pred_depth, gt_depth, min_depth=1e-3, max_depth=80.0):
    """One image of the KITTI Eigen protocol: bilinear resize of the prediction to the ground-truth size, validity
    0.001 < gt < 80 inside the Garg/Eigen crop, median scaling, clamp, metrics -- and the un-scaled metrics next to them.
    Returns dict(ratio, error, abs_error) like KittiEigenEvaluator._single_loss."""
    import cv2
    gh, gw = gt_depth.shape[:2]
    pred = cv2.resize(pred_depth, (gw, gh))
    mask = np.logical_and(gt_depth > min_depth, gt_depth < max_depth)
    y0, y1, x0, x1 = np.array([0.40810811 * gh, 0.99189189 * gh, 0.03594771 * gw, 0.96405229 * gw]).astype(np.int32)
    crop = np.zeros(mask.shape, dtype=bool)
    crop[y0:y1, x0:x1] = True
    mask &= crop
    pred, gt = pred[mask], gt_depth[mask]
    if pred.size == 0:
        raise ValueError("no valid ground-truth pixel inside the evaluation crop")
    ratio = np.median(gt) / np.median(pred)
    scaled = np.clip(pred * ratio, min_depth, max_depth)
    return dict(ratio=ratio, error=compute_errors(gt, scaled), abs_error=compute_errors(gt, np.clip(pred, min_depth, max_depth)))
