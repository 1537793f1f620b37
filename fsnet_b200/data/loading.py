"""Batching utilities with the reference's semantics: dict ``collate_fn`` and ``ConcatDataset``
(vision_base/data/datasets/dataset_utils.py:8-60), ``build_dataloader`` (dataloader_builder.py:5-17)
and the rank-strided infinite-permutation ``TrainingSampler`` (distributed_sampler.py:6-56)."""
import itertools
from typing import Callable, List

import numpy as np
import torch
from torch.utils.data import DataLoader
from torch.utils.data.sampler import Sampler

from ..utils.builder import build


def collate_fn(batch):
    """Stack tensors / ndarrays of the keys shared by every sample; everything else becomes a list."""
    shared = set(batch[0].keys())
    for item in batch[1:]:
        shared &= set(item.keys())
    out = {}
    for key in shared:
        v = batch[0][key]
        if isinstance(v, torch.Tensor):
            out[key] = torch.stack([item[key] for item in batch], dim=0)
        elif isinstance(v, np.ndarray):
            out[key] = torch.stack([torch.from_numpy(item[key]) for item in batch], dim=0)
        else:
            out[key] = [item[key] for item in batch]
    return out


class ConcatDataset(torch.utils.data.Dataset):
    def __init__(self, cfg_list: List[dict], **common_keywords):
        super().__init__()
        self.children = [build(**{**common_keywords, **item}) for item in cfg_list]
        self.seperator = np.cumsum([0] + [len(c) for c in self.children[:-1]])
        self.total_length = int(self.seperator[-1] + len(self.children[-1]))

    def __len__(self):
        return self.total_length

    def __getitem__(self, index):
        child = int(np.searchsorted(self.seperator, index, side="right") - 1)
        return self.children[child][index - int(self.seperator[child])]


class TrainingSampler(Sampler):
    """Every rank walks the same default-seeded permutation and keeps ``indices[rank::world_size]``;
    one pass over ``size`` indices per ``__iter__`` (the generator is never reseeded)."""

    def __init__(self, size: int, rank: int = -1, world_size: int = 1, shuffle: bool = True):
        if not isinstance(size, int):
            raise TypeError(f"TrainingSampler(size=) expects an int. Got type {type(size)}.")
        if size <= 0:
            raise ValueError(f"TrainingSampler(size=) expects a positive int. Got {size}.")
        self._size, self._shuffle, self._rank, self._world_size = size, shuffle, rank, world_size
        self.generator = torch.Generator()

    def __len__(self):
        return self._size

    def __iter__(self):
        idx = torch.randperm(self._size, generator=self.generator).tolist() if self._shuffle else list(range(self._size))
        yield from itertools.islice(idx, max(self._rank, 0), None, self._world_size)


def build_dataloader(dataset, num_workers: int, batch_size: int, collate_fn: Callable, local_rank: int = -1,
                     world_size: int = 1, sampler_cfg: dict = dict(), **kwargs):
    sampler_cfg = dict(sampler_cfg)
    name = sampler_cfg.pop("name", "vision_base.data.dataloader.distributed_sampler.TrainingSampler")
    sampler = build(name, size=len(dataset), rank=local_rank, world_size=world_size, **sampler_cfg)
    return DataLoader(dataset, num_workers=num_workers, batch_size=batch_size, collate_fn=collate_fn, sampler=sampler,
                      drop_last=True, **kwargs)


_UPLOAD_STREAMS = {}


def _upload_stream(device):
    """ONE upload stream per device for the life of the process.  A fresh torch.cuda.Stream per epoch / iterator walks through
    torch's stream pool, and once more streams exist than the GPU has hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default)
    the upload stream shares a queue with the compute stream: copies and training steps serialise, and erratically so (measured
    on B200, tools/diag_e2e.py: 7.2 ms/step for the first iterators, 8-18 ms/step after a dozen)."""
    if device.type != "cuda":
        return None
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _UPLOAD_STREAMS.get(key)
    if st is None:
        st = _UPLOAD_STREAMS[key] = torch.cuda.Stream(device)
    return st


class DevicePrefetcher:
    """Iterates a loader of HOST batches one batch ahead of the consumer: while step k computes, the tensors of batch k+1
    travel from (pinned) host memory to the device on a side stream, so the copy engine works under the training step
    instead of in front of it.  Hand-over is stream-ordered (the consumer's stream waits for the upload's event; the
    caching allocator is told about the consuming stream), there is no host synchronisation.

    The reference uploads inside the training hook (``data[key].cuda()``, base_training_hooks.py:33-37), i.e. copy and
    compute are serialised; a batch that already lives on the device passes through that hook unchanged, so wrapping the
    loader is all it takes:   ``for data in DevicePrefetcher(dataloader): training_hook(data, ...)``.

    ``device_transform`` (optional) is applied to every uploaded batch on the upload stream (device-side augmentation).
    ``device='cpu'`` makes it a plain look-ahead iterator (used by the CPU tests of the ordering logic)."""

    def __init__(self, loader, device=None, depth: int = 1, device_transform=None):
        self.loader, self.depth = loader, max(int(depth), 1)
        self.device_transform = device_transform       # e.g. DeviceAugmentStage: runs on the upload stream, right behind the copy
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch, stream, slot=0):
        if stream is None:
            out = {k: (v.to(self.device) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            return (out if self.device_transform is None else self.device_transform(out)), None
        bufs = self._slots[slot]
        with torch.cuda.stream(stream):
            if self._released[slot] is not None:
                stream.wait_event(self._released[slot])         # the consumer of this slot's previous batch has queued all its reads
            out = {}
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    dst = bufs.get(k)
                    if dst is None or dst.shape != v.shape or dst.dtype != v.dtype:
                        dst = bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    dst.copy_(v, non_blocking=True)
                    out[k] = dst
                else:
                    out[k] = v
            if self.device_transform is not None:
                out = self.device_transform(out)
            event = torch.cuda.Event()
            event.record(stream)
        return out, event

    def __iter__(self):
        """Uploads land in ``depth + 1`` persistent sets of device buffers used round-robin (no allocator traffic per batch, which
        matters once the host runs ahead of the GPU: freshly allocated tensors tied to two streams made the caching allocator
        synchronise).  A yielded batch therefore stays valid until ``depth`` more batches have been requested -- consume it inside
        the loop body, as the training loop does; the hand-back is stream-ordered (an event recorded on the consumer's stream when
        it asks for the next batch)."""
        from collections import deque
        stream = _upload_stream(self.device)
        source = iter(self.loader)
        queue = deque()
        n_slots = self.depth + 1
        if len(getattr(self, "_slots", ())) != n_slots:
            self._slots = [dict() for _ in range(n_slots)]     # kept across epochs: the buffers are allocated once
        self._released = [None] * n_slots
        state = {"next": 0}

        def pull():
            try:
                slot = state["next"] % n_slots
                batch, event = self._upload(next(source), stream, slot)
                queue.append((batch, event, slot))
                state["next"] += 1
            except StopIteration:
                pass

        for _ in range(self.depth):
            pull()
        while queue:
            batch, event, slot = queue.popleft()
            pull()                                         # the next upload is in flight before this batch is consumed
            if event is not None:
                current = torch.cuda.current_stream(self.device)
                current.wait_event(event)
                for v in batch.values():
                    if isinstance(v, torch.Tensor):
                        v.record_stream(current)
            yield batch
            if stream is not None:                         # the consumer is back for more: everything it does with `batch` is queued
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self._released[slot] = ev
