"""Loss / image logging with the reference's interface (vision_base/utils/logger.py:6-93).

One behavioural improvement that stays API compatible: ``LossLogger.update`` keeps the loss tensors on
the device and only converts them when ``log`` is called, so the training step has no host syncs (the
reference calls ``.item()`` on every entry every step)."""
from typing import Dict

import torch


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n: int = 1):
        self.val = val
        self.sum = self.sum + val * n
        self.count += n

    @property
    def avg(self):
        s = self.sum / max(self.count, 1)
        return float(s) if isinstance(s, torch.Tensor) else s


class LogImageStruct(object):
    def __init__(self, data: torch.Tensor, dataformat: str = "NCHW"):
        self.data, self.dataformat = data, dataformat

    def update(self, data: torch.Tensor):
        self.data = data

    def log_images(self, writer, tag: str, *args, **kwargs):
        writer.add_images(tag, self.data, dataformats=self.dataformat, **kwargs)


class LossLogger():
    def __init__(self, recorder, data_split="train"):
        self.recorder, self.data_split = recorder, data_split
        self.reset()

    def reset(self):
        self.loss_stats: Dict[str, AverageMeter] = {}
        self.hm_stats: Dict[str, LogImageStruct] = {}

    def update(self, loss_dict):
        for key, value in loss_dict.items():
            if key not in self.loss_stats:
                self.loss_stats[key] = AverageMeter()
            v = value.detach().mean() if isinstance(value, torch.Tensor) else float(value)
            self.loss_stats[key].update(v)

    def log(self, step):
        for key, meter in self.loss_stats.items():
            self.recorder.add_scalar(key + "/" + self.data_split, meter.avg, step)
        for key, hm in self.hm_stats.items():
            hm.log_images(self.recorder, key + "/" + self.data_split, global_step=step)

    def update_hm(self, feature_map_dict):
        for key, value in feature_map_dict.items():
            if isinstance(value, dict):
                data, fmt = value["data"], value.get("dataformat", "NCHW")
            else:
                data, fmt = value, ("NCHW" if value.dim() == 4 else "CHW")
            if key not in self.hm_stats:
                self.hm_stats[key] = LogImageStruct(data, fmt)
            else:
                self.hm_stats[key].update(data)


def styling_git_info(repo):
    log = ("\n-----------------\n# Git Last Commit\n" + repo.git.log(-1) + "\n\n-----------------\n# Git Diff\n\n")
    return log.replace(" ", "&nbsp;").replace("\n", "  \n") + f"```diff\n{repo.git.diff()}\n```"
