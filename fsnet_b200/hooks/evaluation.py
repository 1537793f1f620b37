"""Evaluation hooks with the reference's names, keywords and call signature ``hook(meta_arch, dataset_val, writer,
global_step, epoch_num)`` (vision_base/pipeline_hooks/evaluation_hooks/base_evaluation_hooks.py:12-48,
monodepth/pipeline_hooks/evaluation_hooks/base_evaluation_hooks.py:19-64, 146-202).

The network call is the validation hook's (``test_run_hook_cfg``); everything here is host-side bookkeeping: undo the
pad-resize (``('image_resize','effective_size')``), bring the prediction back to the size of the original frame, hand it to
the dataset evaluator and average what it returns.  Extra keywords become attributes (``batch_size``, ``num_workers``)."""
from typing import Dict, Optional

import cv2
import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from ..data.loading import collate_fn
from ..utils.builder import build

try:
    from tqdm import tqdm
except ImportError:  # pragma: no cover
    def tqdm(x, **kwargs):
        return x


class BaseEvaluationHook(object):
    """Sample-by-sample: ``dataset_eval.reset()``, then ``dataset_eval.step(index, output, sample)`` per sample, then
    ``dataset_eval(writer, global_step, epoch_num)`` unless the split is 'test'."""

    def __init__(self, test_run_hook_cfg, dataset_eval_cfg, result_path_split: str = "validation", **kwargs):
        self.test_hook = build(**test_run_hook_cfg)
        self.result_path_split = result_path_split
        self.dataset_eval = build(**dataset_eval_cfg)
        for key, value in kwargs.items():
            setattr(self, key, value)

    @torch.no_grad()
    def __call__(self, meta_arch: nn.Module, dataset_val, writer=None, global_step: int = 0, epoch_num: int = 0):
        meta_arch.eval()
        self.dataset_eval.reset()
        for index in tqdm(range(len(dataset_val)), dynamic_ncols=True):
            data = dataset_val[index]
            collated: Dict = collate_fn([data])
            self.dataset_eval.step(index, self.test_hook(collated, meta_arch, global_step, epoch_num), data)
        if self.result_path_split != "test" and self.dataset_eval is not None:
            self.dataset_eval(writer, global_step, epoch_num)


class _DepthEvaluationHook(object):
    DEFAULT_BATCH = 1

    def __init__(self, test_run_hook_cfg, dataset_eval_cfg: Optional[dict] = None, **kwargs):
        self.test_hook = build(**test_run_hook_cfg)
        self.dataset_eval_func = None if dataset_eval_cfg is None else build(**dataset_eval_cfg)
        for key, value in kwargs.items():
            setattr(self, key, value)

    def _predictions(self, meta_arch, dataset_val, global_step, epoch_num):
        """yields (batch, i, depth [h_eff, w_eff] float numpy, (h, w) of the original frame) in dataset order."""
        loader = DataLoader(dataset_val, getattr(self, "batch_size", self.DEFAULT_BATCH), shuffle=False,
                            num_workers=getattr(self, "num_workers", 4), collate_fn=collate_fn)
        for batch in tqdm(loader):
            out = self.test_hook(batch, meta_arch, global_step, epoch_num)
            depth = out["depth"].detach().float().cpu().numpy()
            for i in range(depth.shape[0]):
                h_eff, w_eff = (int(v) for v in batch[("image_resize", "effective_size")][i])
                h, w = batch[("original_image", 0)][i].shape[:2]
                yield batch, i, depth[i, 0, :h_eff, :w_eff], (int(h), int(w))


class KittiEvaluationHook(_DepthEvaluationHook):
    """Eigen protocol: the k-th prediction is compared with the k-th ground-truth map; the up-sampling to the original
    size is done in inverse depth (base_evaluation_hooks.py:55-60)."""

    @torch.no_grad()
    def __call__(self, meta_arch: nn.Module, dataset_val, writer=None, global_step: int = 0, epoch_num: int = 0):
        meta_arch.eval()
        errors, abs_errors = [], []
        for frame_index, (_, _, depth, (h, w)) in enumerate(self._predictions(meta_arch, dataset_val, global_step, epoch_num)):
            res = self.dataset_eval_func.single_call(1 / cv2.resize(1 / depth, (w, h)), frame_index)
            errors.append(res["error"])
            abs_errors.append(res["abs_error"])
        mean_errors, mean_abs_errors = np.array(errors).mean(0), np.array(abs_errors).mean(0)
        self.dataset_eval_func.log(writer, mean_errors, mean_abs_errors, global_step=global_step, epoch_num=epoch_num)
        return dict(error=mean_errors, abs_error=mean_abs_errors)


class FastNuscEvaluationHook(_DepthEvaluationHook):
    """Per-camera means keyed by ``camera_type``; ground truth looked up by ``('filename', 0)``; depth is resized directly;
    frames without usable LiDAR points (the evaluator raises ValueError) are skipped (base_evaluation_hooks.py:146-202)."""
    DEFAULT_BATCH = 16

    @torch.no_grad()
    def __call__(self, meta_arch: nn.Module, dataset_val, writer=None, global_step: int = 0, epoch_num: int = 0):
        import warnings
        meta_arch.eval()
        errors, abs_errors = {}, {}
        for batch, i, depth, (h, w) in self._predictions(meta_arch, dataset_val, global_step, epoch_num):
            cam = batch["camera_type"][i]
            errors.setdefault(cam, [])
            abs_errors.setdefault(cam, [])
            if self.dataset_eval_func is None:
                continue
            filename = batch[("filename", 0)][i]
            try:
                res = self.dataset_eval_func.single_call(cv2.resize(depth, (w, h)), filename)
            except ValueError:
                warnings.warn(f"image at sample {filename}  has no usable points")
                continue
            errors[cam].append(res["error"])
            abs_errors[cam].append(res["abs_error"])
        per_cam, per_cam_abs = [], []
        for cam in errors:
            mean_errors, mean_abs_errors = np.array(errors[cam]).mean(0), np.array(abs_errors[cam]).mean(0)
            self.dataset_eval_func.log(writer, cam, mean_errors, mean_abs_errors, global_step=global_step, epoch_num=epoch_num)
            per_cam.append(mean_errors)
            per_cam_abs.append(mean_abs_errors)
        all_mean, all_mean_abs = np.array(per_cam).mean(0), np.array(per_cam_abs).mean(0)
        self.dataset_eval_func.log(writer, "all mean", all_mean, all_mean_abs, global_step=global_step, epoch_num=epoch_num)
        return dict(error=all_mean, abs_error=all_mean_abs)
