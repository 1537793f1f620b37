"""Optimiser / scheduler factories and checkpoint helpers with the reference's names and behaviour
(vision_base/networks/optimizers/{optimizers,schedulers}.py, vision_base/networks/utils/utils.py:3-19)."""
import torch
import torch.nn as nn
import torch.optim as optim


def build_optimizer(model: nn.Module, name, **kwargs):
    table = {"sgd": optim.SGD, "adam": optim.Adam, "adamw": optim.AdamW}
    if name.lower() not in table:
        raise NotImplementedError(name)
    return table[name.lower()](model.parameters(), **kwargs)


class PolyLR(optim.lr_scheduler._LRScheduler):
    def __init__(self, optimizer, gamma=0.9, n_iteration=-1):
        self.step_size, self.gamma = n_iteration, gamma
        super().__init__(optimizer)

    def get_lr(self):
        decay = max(0.0, 1 - self._step_count / float(self.step_size)) ** self.gamma
        return [base_lr * decay for base_lr in self.base_lrs]


def build_scheduler(optimizer, name=None, **kwargs):
    if name is None:
        return optim.lr_scheduler.ExponentialLR(optimizer, 1.0)
    table = {"steplr": optim.lr_scheduler.StepLR, "multisteplr": optim.lr_scheduler.MultiStepLR,
             "exponentiallr": optim.lr_scheduler.ExponentialLR, "cosineannealinglr": optim.lr_scheduler.CosineAnnealingLR,
             "polylr": PolyLR}
    if name.lower() not in table:
        raise NotImplementedError(name)
    return table[name.lower()](optimizer, **kwargs)


def _unwrap(model):
    return model.module if (torch.distributed.is_available() and torch.distributed.is_initialized() and hasattr(model, "module")) else model


def save_models(path, model, optimizer=None):
    """{'model_state_dict', 'optimizer_state_dict'} with the DDP wrapper stripped."""
    state = {"model_state_dict": _unwrap(model).state_dict()}
    if optimizer is not None:
        state["optimizer_state_dict"] = optimizer.state_dict()
    torch.save(state, path)


def load_models(path, model, optimizer=None, map_location=None, strict=False):
    state = torch.load(path, map_location=map_location)
    _unwrap(model).load_state_dict(state["model_state_dict"], strict=strict)
    if optimizer is not None and "optimizer_state_dict" in state:
        optimizer.load_state_dict(state["optimizer_state_dict"])
