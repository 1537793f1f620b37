"""Export a stage-1 training checkpoint as the teacher of the distillation stage (reference CLI:
``python monodepth/transform_teacher.py <checkpoint> <teacher.pth>``, monodepth/transform_teacher.py:6-25).

Keeps the depth encoder (``depth_backbone.*``) and the depth decoder (``head.depth_decoder.*`` -> ``depth_decoder.*``);
the pose head and everything else is dropped.  The result is the flat state dict ``DistillWPoseMeta`` loads into
``teacher_net``."""
import sys
from collections import OrderedDict

import torch


def transform_teacher_model(src_model_path: str, tar_model_path: str):
    state = torch.load(src_model_path, map_location="cpu")["model_state_dict"]
    teacher = OrderedDict()
    for key, value in state.items():
        key = key[len("module."):] if key.startswith("module.") else key          # checkpoints written from a DDP wrapper
        if key.startswith("depth_backbone"):
            teacher[key] = value
        elif key.startswith("head.depth_decoder"):
            teacher[key[len("head."):]] = value
    torch.save(teacher, tar_model_path)
    return teacher


if __name__ == "__main__":
    try:
        from fire import Fire
        Fire(transform_teacher_model)
    except ImportError:
        transform_teacher_model(*sys.argv[1:3])
