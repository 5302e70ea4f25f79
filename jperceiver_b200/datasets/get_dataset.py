"""``mono.datasets.get_dataset.get_dataset`` (get_dataset.py:9-44).

The reference's dataset classes are file readers (pykitti / Argoverse trees, split files) and outside this path's scope
(SURVEY.md §8); they are plain map-style torch datasets with a ``flag`` array and work with ``build_dataloader`` as they are.
What this module provides is the dispatch by ``cfg['name']`` and the one dataset that needs no files: ``name='synthetic'``
yields snippets with exactly the keys / shapes / dtypes ``MonoDataset.__getitem__`` emits (``jperceiver_b200.synthetic``), so the
reference's ``train.py`` can drive a complete run without KITTI on disk."""
from __future__ import annotations

import numpy as np
import torch

from .. import synthetic


class SyntheticSnippets(torch.utils.data.Dataset):
    """``len`` = ``cfg.num_samples`` (default 64); sample ``i`` is ``synthetic.make_batch(..., B=1, seed=seed0 + i)`` without
    the batch dimension.  ``flag`` is all zeros as in ``mono_dataset.py`` (one aspect-ratio group)."""

    def __init__(self, cfg, training=True):
        self.opt = dict(height=int(cfg["height"]), width=int(cfg["width"]), frame_ids=list(cfg["frame_ids"]) if training else [0],
                        occ_map_size=int(cfg.get("occ_map_size", 256)), split=cfg.get("split", "odometry"))
        self.training = bool(training)
        self.n = int(cfg.get("num_samples", 64))
        self.seed0 = int(cfg.get("seed", 1024)) + (0 if training else 1 << 20)
        self.flag = np.zeros(self.n, dtype=np.int64)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        b = synthetic.make_batch(self.opt, 1, seed=self.seed0 + int(i))
        d = {k: v[0] for k, v in b.items()}
        if not self.training:      # validation items carry a ground-truth depth frame (0 = no LiDAR return), kitti_dataset.py:66-77
            g = torch.Generator().manual_seed(self.seed0 + int(i))
            h, w = d[("color", 0, -1)].shape[-2:]
            depth = 2.0 + 60.0 * torch.rand(h, w, generator=g)
            d["gt_depth"] = torch.where(torch.rand(h, w, generator=g) < 0.2, depth, torch.zeros(()))
        return d


def get_dataset(cfg, training=True):
    name = cfg["name"]
    if name == "synthetic":
        return SyntheticSnippets(cfg, training)
    raise NotImplementedError(
        "dataset %r: the reference's file readers (mono/datasets/*_dataset.py) are outside the scope of this package — build the "
        "dataset with the reference's own class and pass it to train_mono / build_dataloader, or use name='synthetic'" % (name,))
