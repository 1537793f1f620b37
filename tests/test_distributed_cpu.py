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


# ----------------------------------------------------------------------------------------------------------------------
# The N>1 DATA PATH itself: a SyncBN training step through the executor on two ranks (gloo), kernels under the SIMT emulator
# ----------------------------------------------------------------------------------------------------------------------
def _syncbn_worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _pytest.monkeypatch import MonkeyPatch
    from host_emulation import fixture
    patch = MonkeyPatch()
    fixture.install(patch)
    try:
        from oracle import fsnet_oracle as O
        from helpers import build_model
        from fsnet_b200.hooks.training import BaseTrainingHook
        from fsnet_b200.networks import ops
        ops.set_backend("tc")
        from fsnet_b200 import engine
        engine.Tape.bucketed_allreduce = True      # the hook's setting for models that are not DDP-wrapped: buckets reduced during backward
        posenet = os.environ.get("FSNET_DIST_POSENET") == "1"        # opt-in variant: depth net + PoseNet (two executor tapes per step)
        topo = O.Topology(height=32, width=64, posenet=posenet, overlapped_mask=not posenet)
        B = 2 * world
        data = O.synthetic_batch(B, topo.height, topo.width, 77, topo.frame_ids)
        data.pop("patched_mask")                    # equal loss normalisers on every rank: the mean of rank losses is the global loss
        noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
        meta = dict(is_training=True, epoch_num=0, global_step=0)

        def run(model, lo, hi):
            model.head.tie_break_noise = {s: n[lo:hi] for s, n in noise.items()}
            shard = {k: (v[lo:hi] if torch.is_tensor(v) else v) for k, v in data.items()}
            ret = model(shard, meta)
            ret["loss"].mean().backward()
            return float(ret["loss"].detach())

        # data parallel: SyncBatchNorm statistics all-reduced inside the executor, flat gradient all-reduce afterwards
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(build_model(topo))
        loss = run(model, 2 * rank, 2 * rank + 2)
        BaseTrainingHook.sync_gradients(model)
        losses = [None] * world
        dist.all_gather_object(losses, loss)
        result = None
        engine.Tape.bucketed_allreduce = False     # the reference run below is one process on the whole batch
        if rank == 0:
            single = build_model(topo)              # plain BatchNorm, whole batch, one process
            loss_single = run(single, 0, B)
            worst = 0.0
            ref = dict(single.named_parameters())
            for k, p in model.named_parameters():
                g, r = p.grad.double(), ref[k].grad.double()
                if float(r.norm()) > 1e-7:
                    worst = max(worst, float((g - r).norm() / r.norm()))
            stats = max(float((a - b).abs().max()) for (_, a), (_, b) in zip(model.named_buffers(), single.named_buffers())
                        if a.is_floating_point())
            result = (sum(losses) / world, loss_single, worst, stats)
        out.put(result)
    finally:
        patch.undo()
        dist.destroy_process_group()


def test_syncbn_training_step_world2_matches_single_process():
    """Two ranks x 2 samples with SyncBatchNorm (forward statistics and the two backward sums all-reduced inside the executor,
    engine.py) + the hook's flat gradient all-reduce == one process x 4 samples with plain BatchNorm: loss, every parameter
    gradient, and the running statistics.  The kernels run under the SIMT emulator (tests/host_emulation), the collectives over gloo.
    (scripts/train.py's DistributedDataParallel wrapping can not be exercised here: torch refuses SyncBatchNorm in CPU modules.)"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from host_emulation import emulate
    emulate.build_simt_library()                  # compiled once here, found in the cache by both ranks
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    import time
    results, deadline = [], time.time() + 600
    while len(results) < len(procs):
        try:
            results.append(q.get(timeout=2))
        except queue.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a rank died: " + str([p.exitcode for p in procs])
            assert time.time() < deadline, "timed out"
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    mean_loss, loss_single, worst, stats = next(r for r in results if r is not None)
    assert abs(mean_loss - loss_single) <= 1e-5 * abs(loss_single), (mean_loss, loss_single)
    assert worst < 2e-2, worst                    # bf16 operands in the gradient convolutions; typical 1e-3
    assert stats < 1e-5, stats                    # running mean / var updated from the GLOBAL batch statistics
