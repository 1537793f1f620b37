"""Evaluation launcher with the reference's CLI (scripts/test.py:11-54):

    python scripts/test.py --config=configs/kitti_wpose_files.py --checkpoint_path=<...>.pth [--split_to_test=validation] [--a.b=value ...]

Builds the dataset of the chosen split and the model from the config, loads the checkpoint (non-strict), and runs
``cfg.trainer.evaluate_hook`` over it.  ``fire`` is used when importable; otherwise the ``--key=value`` parser of train.py.
"""
import ast
import sys

from _path_init import manage_package_logging  # noqa: F401  (also fixes sys.path)
import torch

from vision_base.utils.builder import build
from vision_base.utils.utils import cfg_from_file, update_cfg
from vision_base.networks.utils.utils import load_models


def main(config="config/config.py", gpu=0, checkpoint_path="retinanet_79.pth", split_to_test="validation", **kwargs):
    cfg = cfg_from_file(config)
    cfg = update_cfg(cfg, **kwargs)
    cfg.trainer.gpu = gpu
    torch.cuda.set_device(cfg.trainer.gpu)
    manage_package_logging()

    split_cfg = {"training": "train_dataset", "test": "test_dataset"}.get(split_to_test, "val_dataset")
    dataset = build(**cfg[split_cfg])

    meta_arch = build(**cfg.meta_arch).cuda()
    load_models(checkpoint_path, meta_arch, map_location=f"cuda:{gpu}", strict=False)
    meta_arch.eval()

    if "evaluate_hook" not in cfg.trainer:
        raise KeyError("evaluate_hook not found in Config")
    evaluate_hook = build(result_path_split="validation", **cfg.trainer.evaluate_hook)
    print("Found evaluate function")
    result = evaluate_hook(meta_arch, dataset)
    print("finish")
    return result


def _cli():
    try:
        from fire import Fire
        return Fire(main)
    except ImportError:
        kwargs = {}
        for arg in sys.argv[1:]:
            assert arg.startswith("--") and "=" in arg, f"expected --key=value, got {arg}"
            k, v = arg[2:].split("=", 1)
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
            kwargs[k.replace("-", "_") if k in ("checkpoint-path", "split-to-test") else k] = v
        return main(**kwargs)


if __name__ == "__main__":
    _cli()
