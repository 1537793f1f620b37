"""Training launcher with the reference's CLI (scripts/train.py:21-214):

    python scripts/train.py --config=configs/kitti_wpose_synthetic.py --experiment_name=run [--a.b.c=value ...]
    torchrun --nproc-per-node N scripts/train.py --config=... --world_size=N       (LOCAL_RANK from the env)

``fire`` is used when importable; otherwise a small ``--key=value`` parser with the same conventions.
"""
import ast
import os
import shutil
import sys

from _path_init import manage_package_logging  # noqa: F401  (also fixes sys.path)
import torch
from easydict import EasyDict

from vision_base.utils.builder import build
from vision_base.utils.utils import get_num_parameters, cfg_from_file, set_random_seed, update_cfg
from vision_base.utils.timer import Timer
from vision_base.utils.logger import LossLogger
from vision_base.data.datasets.dataset_utils import collate_fn
from vision_base.data.dataloader import build_dataloader
from vision_base.networks.optimizers import optimizers, schedulers
from vision_base.networks.utils.utils import save_models, load_models


class _NullWriter:
    def __getattr__(self, name):
        return lambda *a, **k: None


def build_train_loader(cfg, dataset_train, local_rank, world_size, device):
    """The reference's loader (scripts/train.py:77-81) plus two opt-in stages in front of the training hook:
    the upload of batch k+1 runs on a side stream while step k computes (default; FSNET_PREFETCH=0 = the reference: upload inside the hook, serialised with
    the step); a dataset whose augmentation is fsnet_b200.data.device_augment.DeviceAugmentation ships uint8 frames + drawn
    parameters, and the pixel work runs on the GPU right behind that upload."""
    from fsnet_b200.data.device_augment import device_augment_collate, find_device_stage
    device_stage = find_device_stage(dataset_train)
    prefetch = bool(int(os.environ.get("FSNET_PREFETCH", "1"))) or device_stage is not None
    loader = build_dataloader(dataset_train, num_workers=cfg.data.num_workers, batch_size=cfg.data.batch_size,
                              collate_fn=collate_fn if device_stage is None else device_augment_collate,
                              local_rank=local_rank, world_size=world_size, sampler_cfg=getattr(cfg.data, "sampler", dict()),
                              pin_memory=prefetch)
    if prefetch:
        from fsnet_b200.data.loading import DevicePrefetcher
        loader = DevicePrefetcher(loader, device, device_transform=device_stage)
    return loader


def main(config="configs/config.py", experiment_name="default", world_size=1, local_rank=-1, **kwargs):
    cfg = cfg_from_file(config)
    cfg = update_cfg(cfg, **kwargs)
    if local_rank < 0 and world_size > 1 and "LOCAL_RANK" in os.environ:
        local_rank = int(os.environ["LOCAL_RANK"])
    cfg.dist = EasyDict(world_size=world_size, local_rank=local_rank)
    is_distributed = local_rank >= 0
    is_logging = local_rank <= 0

    writer = None
    if is_logging:
        recorder_dir = os.path.join(cfg.path.log_path, experiment_name + "config=" + os.path.basename(config))
        if os.path.isdir(recorder_dir):
            shutil.rmtree(recorder_dir, ignore_errors=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(recorder_dir)
            import pprint
            writer.add_text("config.py", pprint.pformat(cfg).replace(" ", "&nbsp;").replace("\n", "  \n"))
        except Exception:  # noqa: BLE001 - tensorboard is optional
            writer = _NullWriter()

    if is_distributed:
        cfg.trainer.gpu = local_rank
    gpu = min(cfg.trainer.gpu, torch.cuda.device_count() - 1)
    torch.backends.cudnn.benchmark = getattr(cfg.trainer, "cudnn", False)
    set_random_seed(123)
    torch.cuda.set_device(gpu)
    if is_distributed:
        torch.distributed.init_process_group(backend="nccl", init_method="env://")

    if "precompute_hook" in cfg:
        build(**cfg.precompute_hook)()

    dataset_train = build(**cfg.train_dataset)
    dataset_val = build(**cfg.val_dataset) if "val_dataset" in cfg else None
    dataloader_train = build_train_loader(cfg, dataset_train, local_rank, world_size, torch.device("cuda", gpu))

    meta_arch = build(**cfg.meta_arch)
    from vision_base.networks.models.meta_archs.base_meta import BaseMetaArch
    assert isinstance(meta_arch, BaseMetaArch)
    if is_distributed:
        meta_arch = torch.nn.SyncBatchNorm.convert_sync_batchnorm(meta_arch)
        meta_arch = torch.nn.parallel.DistributedDataParallel(meta_arch.cuda(), device_ids=[gpu], output_device=gpu)
    else:
        meta_arch = meta_arch.cuda()
    meta_arch.train()
    if is_logging:
        print(f"number of trained parameters of the model: {get_num_parameters(meta_arch)}")

    optimizer = optimizers.build_optimizer(meta_arch, **cfg.optimizer)
    scheduler_config = getattr(cfg, "scheduler", None) or {}
    scheduler = schedulers.build_scheduler(optimizer, **scheduler_config)
    is_iter_based = scheduler_config.get("is_iter_based", False)
    training_loss_logger = LossLogger(writer, "train") if is_logging else None
    manage_package_logging()

    old_checkpoint = getattr(cfg.path, "pretrained_checkpoint", None)
    if old_checkpoint is not None:
        load_models(old_checkpoint, meta_arch.module if is_distributed else meta_arch, optimizer, map_location=f"cuda:{gpu}")

    if "training_hook" not in cfg.trainer:
        raise KeyError
    training_hook = build(**cfg.trainer.training_hook)
    from vision_base.pipeline_hooks.train_val_hooks.base_training_hooks import BaseTrainingHook
    assert isinstance(training_hook, BaseTrainingHook)
    evaluate_hook = build(result_path_split="validation", **cfg.trainer.evaluate_hook) if "evaluate_hook" in cfg.trainer else None

    timer = Timer()
    print(f"Num training images: {len(dataset_train)}")
    global_step = 0
    max_steps = getattr(cfg.trainer, "max_steps", None)
    for epoch_num in range(cfg.trainer.max_epochs):
        meta_arch.train()
        if training_loss_logger:
            training_loss_logger.reset()
        for iter_num, data in enumerate(dataloader_train):
            training_hook(data, meta_arch, optimizer, writer, training_loss_logger, global_step, epoch_num)
            global_step += 1
            if is_iter_based:
                scheduler.step()
            if is_logging and global_step % cfg.trainer.disp_iter == 0 and "total_loss" in training_loss_logger.loss_stats:
                log_str = "Epoch: {} | Iteration: {}  | Running loss: {:1.5f} | eta:{}".format(
                    epoch_num, iter_num, training_loss_logger.loss_stats["total_loss"].avg,
                    timer.compute_eta(global_step, len(dataloader_train) * cfg.trainer.max_epochs / world_size))
                print(log_str, end="\r")
                writer.add_text("training_log/train", log_str, global_step)
                training_loss_logger.log(global_step)
            if max_steps is not None and global_step >= max_steps:
                break
        if not is_iter_based:
            scheduler.step()
        if is_logging:
            save_models(os.path.join(cfg.path.checkpoint_path, f"{cfg.meta_arch.name}_latest.pth"), meta_arch, optimizer)
            if (epoch_num + 1) % cfg.trainer.save_iter == 0:
                save_models(os.path.join(cfg.path.checkpoint_path, f"{cfg.meta_arch.name}_{epoch_num}.pth"), meta_arch, optimizer)
        if is_logging and evaluate_hook is not None and cfg.trainer.test_iter > 0 and (epoch_num + 1) % cfg.trainer.test_iter == 0:
            evaluate_hook(meta_arch.module if is_distributed else meta_arch, dataset_val, writer, epoch_num, epoch_num)
        if is_distributed:
            torch.distributed.barrier()
        if is_logging:
            writer.flush()
        if max_steps is not None and global_step >= max_steps:
            break
    if is_logging:
        print(f"\nfinished {global_step} steps")
    return global_step


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
            kwargs[k.replace("-", "_") if k in ("local-rank", "world-size", "experiment-name") else k] = v
        return main(**kwargs)


if __name__ == "__main__":
    _cli()
