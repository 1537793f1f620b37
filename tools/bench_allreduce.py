"""NCCL all-reduce time of the gradient-sized buffers (torchrun, N ranks): what the flat gradient exchange of a step has to cost at least."""
import os
import torch
import torch.distributed as dist

dist.init_process_group("nccl")
rank = dist.get_rank()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
for numel, name in ((11_700_000, "parameter gradients (45 MB fp32)"), (14_300_000, "accumulator pool (57 MB)"), (10_000, "BatchNorm / bias gradients (40 KB)")):
    t = torch.randn(numel, device="cuda")
    for op in (dist.ReduceOp.AVG,):
        for _ in range(5):
            dist.all_reduce(t, op=op)
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(t, op=op)
        b.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"{name}: {a.elapsed_time(b) / 20 * 1e3:.1f} us per all-reduce at {dist.get_world_size()} ranks", flush=True)
