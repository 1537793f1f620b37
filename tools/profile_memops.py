"""Which non-kernel device operations (memcpy / memset nodes) does one training step contain?  ncu launch lists only show kernels."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from fsnet_b200.data.synthetic import make_batch
from vision_base.utils.builder import build
from vision_base.utils.utils import cfg_from_file, set_random_seed

cfg = cfg_from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "kitti_wpose_synthetic.py"))
set_random_seed(123)
model = build(**cfg.meta_arch).cuda().train()
from vision_base.networks.optimizers.optimizers import build_optimizer
opt = build_optimizer(model, **cfg.optimizer)
hook = build(**dict(cfg.trainer.training_hook, cuda_graph=False))
data = make_batch(12, 192, 640, device="cuda")
for i in range(3):
    hook(dict(data), model, opt, None, None, i, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    hook(dict(data), model, opt, None, None, 3, 0)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name
        if name.lower().startswith(("memcpy", "memset")):
            k = name.split("(")[0].strip() + (" " + name[name.index("("):] if "(" in name else "")
            agg[k][0] += 1
            agg[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
kern = sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and not e.name.lower().startswith(("memcpy", "memset")))
print(f"kernels: {kern}")
for k, (n, t, _) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k}: {n} operations, {t:.0f} us")
