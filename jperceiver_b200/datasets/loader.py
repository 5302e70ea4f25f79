"""Snippet partitioning across ranks and the batch loader (mirror of ``mono.datasets.loader``).

Reference: ``/root/reference/mono/datasets/loader/sampler.py`` (``DistributedSampler`` :17-41, ``GroupSampler`` :44-81,
``DistributedGroupSampler`` :84-157) and ``build_loader.py:19-56``.  The path is data parallel over snippets: every epoch the
snippet indices of each aspect-ratio group (``dataset.flag``) are permuted with an epoch-seeded generator, padded to a multiple
of ``samples_per_gpu * world``, cut into per-step chunks of ``samples_per_gpu``, the chunks are permuted, and rank ``r`` takes
the ``r``-th contiguous slice.  This is index work: the sequences below are **identical** to the reference's for the same
torch / numpy (``tests/golden/kat_sampler.json`` holds sequences produced by the reference's own classes).

Design differences (B200-first, same results): the whole job's plan for an epoch is one array computed once (``plan``) — every
rank derives the same plan from the epoch alone, so there is no exchange — and a rank's iterator is a view of its row; the
loader pins its staging buffers (the reference passes ``pin_memory=False`` and copies pageable memory) so the H2D copy of the
next batch overlaps the running step (``TrainEngine``'s input pipeline).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Sampler
from torch.utils.data._utils.collate import default_collate


def get_dist_info():
    """(rank, world_size) — what ``mmcv.runner.get_dist_info`` returns (build_loader.py:27)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _pad_wrap(seq: np.ndarray, total: int) -> np.ndarray:
    """``seq + seq[:total - len(seq)]`` (sampler.py:33,139): a single wrap — a group smaller than the padding trips the same
    assertion as in the reference."""
    out = np.concatenate([seq, seq[:max(total - len(seq), 0)]])
    assert len(out) == total, "group of %d samples cannot be padded to %d by one wrap (sampler.py:142)" % (len(seq), total)
    return out


class DistributedSampler(Sampler):
    """Strided split of an (optionally epoch-shuffled) index list (sampler.py:17-41); used for ``shuffle=False`` loaders."""

    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True):
        r, w = get_dist_info()
        self.dataset = dataset
        self.num_replicas = w if num_replicas is None else num_replicas
        self.rank = r if rank is None else rank
        self.shuffle = shuffle
        self.epoch = 0
        self.num_samples = int(math.ceil(len(dataset) / self.num_replicas))
        self.total_size = self.num_samples * self.num_replicas

    def plan(self, epoch=None) -> np.ndarray:
        """[world, num_samples] indices of every rank for ``epoch``."""
        n = len(self.dataset)
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.epoch if epoch is None else epoch)
            order = torch.randperm(n, generator=g).numpy()
        else:
            order = np.arange(n)
        order = _pad_wrap(order, self.total_size)
        return order.reshape(self.num_samples, self.num_replicas).T      # rank r takes order[r::world]

    def __iter__(self):
        return iter(self.plan()[self.rank].tolist())

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


def _group_sizes(flag):
    flag = np.asarray(flag).astype(np.int64)
    return flag, np.bincount(flag)


class GroupSampler(Sampler):
    """Single-process loader order (sampler.py:44-81): batches never mix ``flag`` groups.  Draws from numpy's global generator
    exactly as the reference does (one ``shuffle`` per non-empty group, then one ``permutation`` of the chunks), so
    ``np.random.seed`` reproduces the reference's sequence."""

    def __init__(self, dataset, samples_per_gpu=1):
        assert hasattr(dataset, "flag")
        self.dataset = dataset
        self.samples_per_gpu = samples_per_gpu
        self.flag, self.group_sizes = _group_sizes(dataset.flag)
        self.num_samples = int(sum(-(-int(s) // samples_per_gpu) * samples_per_gpu for s in self.group_sizes))

    def __iter__(self):
        spg = self.samples_per_gpu
        parts = []
        for grp in np.flatnonzero(self.group_sizes):
            members = np.flatnonzero(self.flag == grp)
            np.random.shuffle(members)
            parts.append(_pad_wrap(members, -(-len(members) // spg) * spg))
        chunks = np.concatenate(parts).reshape(-1, spg)
        chunks = chunks[np.random.permutation(range(len(chunks)))]
        order = torch.from_numpy(chunks.reshape(-1)).long()
        assert len(order) == self.num_samples
        return iter(order)

    def __len__(self):
        return self.num_samples


class DistributedGroupSampler(Sampler):
    """The training partition (sampler.py:84-157): epoch-seeded, group-pure chunks of ``samples_per_gpu``, one contiguous slice
    of the chunk-permuted order per rank.  ``plan(epoch)`` is the whole job's assignment; all ranks compute the same plan."""

    def __init__(self, dataset, samples_per_gpu=1, num_replicas=None, rank=None):
        r, w = get_dist_info()
        assert hasattr(dataset, "flag")
        self.dataset = dataset
        self.samples_per_gpu = samples_per_gpu
        self.num_replicas = w if num_replicas is None else num_replicas
        self.rank = r if rank is None else rank
        self.epoch = 0
        self.flag, self.group_sizes = _group_sizes(dataset.flag)
        per_step = samples_per_gpu * self.num_replicas
        self._padded = [int(math.ceil(int(s) / per_step)) * per_step for s in self.group_sizes]
        self.total_size = int(sum(self._padded))
        self.num_samples = self.total_size // self.num_replicas

    def plan(self, epoch=None) -> np.ndarray:
        """[world, num_samples] snippet indices of every rank for ``epoch`` (default: the epoch set by ``set_epoch``)."""
        g = torch.Generator()
        g.manual_seed(self.epoch if epoch is None else epoch)
        parts = []
        for grp in np.flatnonzero(self.group_sizes):
            members = np.flatnonzero(self.flag == grp)
            members = members[torch.randperm(len(members), generator=g).numpy()]
            parts.append(_pad_wrap(members, self._padded[grp]))
        order = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        assert len(order) == self.total_size
        chunks = order.reshape(-1, self.samples_per_gpu)
        chunks = chunks[torch.randperm(len(chunks), generator=g).numpy()]
        return chunks.reshape(self.num_replicas, self.num_samples)

    def __iter__(self):
        return iter(self.plan()[self.rank].tolist())

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


def collate(batch, samples_per_gpu=1):
    """``mmcv.parallel.collate`` for this path's samples (dicts of tensors / arrays / numbers keyed by tuples, no
    ``DataContainer``): stack along a new batch dimension."""
    return default_collate(batch)


def build_dataloader(dataset, imgs_per_gpu, workers_per_gpu, num_gpus=1, dist=True, **kwargs):
    """``mono.datasets.loader.build_dataloader`` (build_loader.py:19-56): same arguments, same sampler choice, same
    ``drop_last=True``; staging memory is pinned (see the module docstring)."""
    shuffle = kwargs.get("shuffle", True)
    if dist:
        rank, world = get_dist_info()
        if shuffle:
            sampler = DistributedGroupSampler(dataset, imgs_per_gpu, world, rank)
        else:
            sampler = DistributedSampler(dataset, world, rank, shuffle=False)
        batch_size, num_workers = imgs_per_gpu, workers_per_gpu
    else:
        sampler = GroupSampler(dataset, imgs_per_gpu) if shuffle else None
        batch_size, num_workers = num_gpus * imgs_per_gpu, num_gpus * workers_per_gpu
    kwargs = dict(kwargs)
    kwargs.pop("shuffle", None)       # a sampler decides the order (the reference forwards the key and DataLoader rejects it)
    kwargs.setdefault("pin_memory", torch.cuda.is_available())
    if num_workers > 0:
        kwargs.setdefault("persistent_workers", True)
    return DataLoader(dataset, batch_size=batch_size, sampler=sampler, num_workers=num_workers,
                      collate_fn=collate, drop_last=True, **kwargs)
