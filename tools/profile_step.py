"""Three eager training steps at cfg2 (B=12, 192x640) for ncu launch lists / kernel captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fsnet_b200.data.synthetic import make_batch
from vision_base.utils.builder import build
from vision_base.utils.utils import cfg_from_file, set_random_seed

cfg = cfg_from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "kitti_wpose_synthetic.py"))
set_random_seed(123)
model = build(**cfg.meta_arch).cuda().train()
from vision_base.networks.optimizers.optimizers import build_optimizer
opt = build_optimizer(model, **cfg.optimizer)
hook = build(**cfg.trainer.training_hook)
data = make_batch(int(os.environ.get("B", 12)), 192, 640, device="cuda")
for i in range(int(os.environ.get("STEPS", 3))):
    hook(dict(data), model, opt, None, None, i, 0)
torch.cuda.synchronize()
print("done")
