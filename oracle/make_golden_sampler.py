"""Generate ``tests/golden/kat_sampler.json`` by running the REFERENCE's own samplers (authoring container only).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.   Usage:  ``python -m oracle.make_golden_sampler``

``/root/reference/mono/datasets/loader/sampler.py`` is loaded by file path (it needs torch and numpy only); every index sequence
written here comes out of its ``DistributedGroupSampler`` / ``DistributedSampler`` / ``GroupSampler``.  The sequences depend on
``torch.randperm`` / ``numpy.random`` of the installed versions (recorded in the file; the GPU box runs the same image).
"""
from __future__ import annotations

import importlib.util
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference/mono/datasets/loader/sampler.py"


class FlagDataset:
    """What the samplers read of a dataset: its length and the per-sample group flag (mono_dataset.py sets ``flag`` to zeros)."""

    def __init__(self, flag):
        self.flag = np.asarray(flag, dtype=np.int64)

    def __len__(self):
        return len(self.flag)


def flags(kind, n):
    if kind == "zeros":                       # every dataset of the reference: one group
        return np.zeros(n, dtype=np.int64)
    if kind == "two":                         # two aspect-ratio groups, interleaved 2:1
        return (np.arange(n) % 3 == 2).astype(np.int64)
    if kind == "gap":                         # group 1 empty
        return np.where(np.arange(n) % 4 == 0, 2, 0).astype(np.int64)
    raise KeyError(kind)


DGS_CASES = [  # (flag kind, n, samples_per_gpu, world, epochs)
    ("zeros", 37, 4, 1, (0, 3)), ("zeros", 37, 4, 2, (0, 1)), ("zeros", 100, 8, 8, (0, 7)), ("two", 50, 3, 2, (0, 5)),
    ("gap", 41, 2, 4, (2,)), ("zeros", 24, 24, 1, (0,)), ("zeros", 5, 1, 4, (0,)),
]
DS_CASES = [(37, 2, False, 0), (37, 4, True, 3), (8, 8, False, 0)]   # (n, world, shuffle, epoch)
GS_CASES = [("zeros", 37, 4, 11), ("two", 50, 3, 12), ("gap", 41, 2, 13)]   # (flag kind, n, samples_per_gpu, numpy seed)


def main():
    spec = importlib.util.spec_from_file_location("_jref_sampler", REF)
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    out = {"torch": torch.__version__, "numpy": np.__version__, "dgs": [], "ds": [], "gs": []}
    for kind, n, spg, world, epochs in DGS_CASES:
        ds = FlagDataset(flags(kind, n))
        for ep in epochs:
            per_rank = []
            for r in range(world):
                s = S.DistributedGroupSampler(ds, spg, world, r)
                s.set_epoch(ep)
                per_rank.append([int(i) for i in s])
            out["dgs"].append({"flag": kind, "n": n, "spg": spg, "world": world, "epoch": ep, "len": len(s), "indices": per_rank})
    for n, world, shuffle, ep in DS_CASES:
        ds = FlagDataset(np.zeros(n))
        per_rank = []
        for r in range(world):
            s = S.DistributedSampler(ds, world, r, shuffle=shuffle)
            s.set_epoch(ep)
            per_rank.append([int(i) for i in s])
        out["ds"].append({"n": n, "world": world, "shuffle": shuffle, "epoch": ep, "len": len(s), "indices": per_rank})
    for kind, n, spg, seed in GS_CASES:
        ds = FlagDataset(flags(kind, n))
        np.random.seed(seed)
        s = S.GroupSampler(ds, spg)
        out["gs"].append({"flag": kind, "n": n, "spg": spg, "seed": seed, "len": len(s), "indices": [int(i) for i in s]})
    with open(os.path.join(GOLD, "kat_sampler.json"), "w") as f:
        json.dump(out, f)
    print("kat_sampler.json: %d + %d + %d cases" % (len(out["dgs"]), len(out["ds"]), len(out["gs"])))


if __name__ == "__main__":
    main()
