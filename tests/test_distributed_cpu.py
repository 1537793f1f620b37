"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sampler sharding and the hook's flat
gradient all-reduce (the N>1 path of bench.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsnet_b200.hooks.training import BaseTrainingHook
    from fsnet_b200.data.loading import TrainingSampler
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    BaseTrainingHook.sync_gradients(model)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(model.parameters()))
    idx = list(TrainingSampler(11, rank=rank, world_size=world))
    gathered = [None] * world
    dist.all_gather_object(gathered, idx)
    if rank == 0:
        flat = sorted(i for g in gathered for i in g)
        out.put((ok, flat == list(range(11)), [len(g) for g in gathered]))
    else:
        out.put((ok, True, None))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_sampler_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in results), "gradients were not averaged across ranks"
    assert all(r[1] for r in results), "sampler shards do not partition the dataset"
