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
