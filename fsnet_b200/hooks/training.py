"""Training / validation hooks (reference: vision_base/pipeline_hooks/train_val_hooks/
base_training_hooks.py:9-49 and base_validation_hooks.py:5-28)."""
from typing import Dict, List, Optional

import torch
import torch.nn as nn


class BaseTrainingHook(object):
    """One optimisation step: zero_grad, host->device, forward, ``loss.mean().backward()``,
    ``clip_grad_norm_``, ``optimizer.step()`` -- same order and semantics as the reference."""

    def __init__(self, tensor_keys: Optional[List[str]] = None, clip_gradients: Optional[float] = None, **kwargs):
        self.tensor_keys = tensor_keys
        self.clip_gradients = clip_gradients

    def __call__(self, data: Dict, meta_arch: nn.Module, optimizer, writer=None, training_loss_logger=None,
                 global_step: int = 0, epoch_num: int = 0):
        optimizer.zero_grad()
        for key in data:
            if isinstance(data[key], torch.Tensor):
                if self.tensor_keys is None or key in self.tensor_keys:
                    data[key] = data[key].cuda(non_blocking=True).contiguous()
        meta = dict(epoch_num=epoch_num, global_step=global_step, is_training=True)
        output: dict = meta_arch(data, meta)
        if training_loss_logger is not None:
            training_loss_logger.update(output["loss_dict"])
            training_loss_logger.update_hm(output.get("hm", dict()))
        output["loss"].mean().backward()
        if self.clip_gradients is not None:
            torch.nn.utils.clip_grad_norm_(meta_arch.parameters(), self.clip_gradients)
        optimizer.step()
        return output


class BaseValidationHook(object):
    """Inference call used by the evaluation hooks (base_validation_hooks.py:5-28)."""

    def __init__(self, tensor_keys: Optional[List[str]] = None, **kwargs):
        self.tensor_keys = tensor_keys

    def __call__(self, data: Dict, meta_arch: nn.Module, global_step: int = 0, epoch_num: int = 0) -> Dict:
        for key in data:
            if isinstance(data[key], torch.Tensor):
                if self.tensor_keys is None or key in self.tensor_keys:
                    data[key] = data[key].cuda().contiguous()
        return meta_arch(data, dict(epoch_num=epoch_num, global_step=global_step, is_training=False))
