"""Config loading / overriding and dotted-name resolution.

Same behaviour as the reference's vision_base/utils/utils.py:38-169 (``cfg_from_file``,
``update_cfg``, ``find_object``) -- configs are Python files defining ``cfg = EasyDict()``.
"""
import importlib
import importlib.util
import os
import random
import uuid

import numpy as np
import torch
from easydict import EasyDict


def get_num_parameters(model) -> int:
    module = getattr(model, "module", model)
    return sum(p.numel() for p in module.parameters() if p.requires_grad)


def set_random_seed(seed: int, deterministic: bool = False) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def cfg_from_file(cfg_filename: str) -> EasyDict:
    """Execute a ``.py`` config and return its ``cfg`` (must be an EasyDict), utils.py:38-53."""
    assert cfg_filename.endswith(".py"), "config files must end in .py"
    name = "_fsnet_cfg_" + uuid.uuid4().hex
    spec = importlib.util.spec_from_file_location(name, os.path.abspath(cfg_filename))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    cfg = getattr(module, "cfg")
    assert isinstance(cfg, EasyDict), "config must define `cfg = EasyDict()`"
    return cfg


def update_cfg(cfg: EasyDict, **kwargs) -> EasyDict:
    """Dotted-key overrides (``--a.b.c=value`` on the CLI), utils.py:82-113: intermediate nodes that
    are missing or not dicts are replaced by fresh EasyDicts."""
    for dotted, value in kwargs.items():
        node = cfg
        parts = dotted.split(".")
        for key in parts[:-1]:
            if not (key in node and isinstance(node[key], dict)):
                node[key] = EasyDict()
            node = node[key]
        node[parts[-1]] = value
    return cfg


def find_object(object_string: str):
    """Longest importable module prefix, then getattr down the rest (utils.py:127-169).
    Raises ModuleNotFoundError carrying every attempt's error when nothing resolves."""
    parts = object_string.split(".")
    traces = []
    for i in range(len(parts), 0, -1):
        prefix = ".".join(parts[:i])
        try:
            obj = importlib.import_module(prefix)
            for attr in parts[i:]:
                obj = getattr(obj, attr)
            return obj
        except Exception as e:  # noqa: BLE001 - mirror the reference: any failure moves on to a shorter prefix
            traces.append(f"{prefix} : {e} ")
    raise ModuleNotFoundError(f"{object_string} not imported, error traces: \n" + "\n".join(traces))
