"""e2e loop variants at cfg2a: which part of the host loop costs what (prefetch depth, loss read-back style)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsnet_b200.data.synthetic import make_batch
from fsnet_b200.data.loading import DevicePrefetcher
from vision_base.utils.builder import build
from vision_base.utils.utils import cfg_from_file, set_random_seed
from vision_base.networks.optimizers.optimizers import build_optimizer
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = cfg_from_file(os.path.join(REPO, "configs", "kitti_wpose_synthetic.py"))
set_random_seed(123)
model = build(**cfg.meta_arch).cuda().train()
opt = build_optimizer(model, **cfg.optimizer)
hook = build(**dict(cfg.trainer.training_hook, cuda_graph=True))
host = make_batch(12, 192, 640, seed=1234)
pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
dev = torch.device("cuda")
resident = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
for i in range(6):
    hook(dict(resident), model, opt, None, None, i, 0)
torch.cuda.synchronize()

def run(name, n, depth, mode):
    slots = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    src = (dict(pinned) for _ in range(n))
    it = DevicePrefetcher(src, dev, depth=depth) if depth > 0 else src
    for i, data in enumerate(it):
        out = hook(data, model, opt, None, None, i, 0)
        if mode == "item":
            out["loss"].item()
        elif mode == "async":
            if i > 0:
                evs[(i - 1) % 2].synchronize()
            slots[i % 2].copy_(out["loss"].detach().reshape(1).double(), non_blocking=True)
            evs[i % 2].record()
        elif mode == "async2":     # lag of two steps
            if i > 1:
                evs[i % 2].synchronize()
            slots[i % 2].copy_(out["loss"].detach().reshape(1).double(), non_blocking=True)
            evs[i % 2].record()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"{name:40s} {ms:7.3f} ms/step", flush=True)

for rep in range(2):
    run("resident, no read", 30, -1, "none") if False else None
    run("prefetch1 + item()", 30, 1, "item")
    run("prefetch1 + async(lag 1)", 30, 1, "async")
    run("prefetch1 + async(lag 2)", 30, 1, "async2")
    run("prefetch2 + async(lag 1)", 30, 2, "async")
    run("prefetch1 + no read", 30, 1, "none")
    run("no prefetch + item()", 30, 0, "item")
    run("no prefetch + no read", 30, 0, "none")
