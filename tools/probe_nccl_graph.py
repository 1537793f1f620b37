"""Can torch.distributed NCCL all_reduce be captured into a CUDA graph on this stack? (2-GPU probe)"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", lr))
x = torch.ones(1024, device="cuda", dtype=torch.float64) * (rank + 1)
y = torch.ones(1 << 20, device="cuda") * (rank + 1)
for _ in range(3):
    dist.all_reduce(x); dist.all_reduce(y)
torch.cuda.synchronize()
print(rank, "eager ok", float(x[0]), flush=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        dist.all_reduce(x)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
x.fill_(rank + 1)
with torch.cuda.graph(g):
    dist.all_reduce(x)
    z = x * 2
    dist.all_reduce(y)
print(rank, "captured", flush=True)
for i in range(3):
    x.fill_(rank + 1); y.fill_(1.0)
    g.replay()
    torch.cuda.synchronize()
    print(rank, "replay", i, float(x[0]), float(z[0]), float(y[0]), flush=True)
dist.destroy_process_group()
