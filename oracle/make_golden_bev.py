"""Golden vectors for the BEV head-loss variants, from the REAL reference (authoring container only).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.   Usage:  ``python -m oracle.make_golden_bev``

Runs the reference's own ``compute_topview_loss`` / ``compute_topview_lossB`` (mono/model/mono_baseline/net.py:554-617) with
every ``loss_type`` in {iou, dice, tversky, focal} and ``loss_sum`` in {1, 2, 3} on the KAT7 inputs of SURVEY.md §8c (logits
from the index pattern at 256x256, B=2, rectangular labels) and writes ``tests/golden/kat_bev_variants.npz``:
``<loss_type>_s<loss_sum>_w<5|15>`` -> float64 loss value, plus the input gradient checksum ``..._gsum`` (sum |dL/dlogits|)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader as R  # noqa: E402
from oracle.make_golden import GOLD, bare_baseline, pat  # noqa: E402


def main():
    net = R.load("registered")
    out = {}
    with R.cpu_cuda_identity():
        big = torch.cat([4 * pat((2, 1, 256, 256), 9) - 2, 4 * pat((2, 1, 256, 256), 10) - 2], 1)
        lab = torch.zeros(2, 1, 256, 256)
        lab[:, :, 64:176, 48:144] = 1
        lab[1, :, 200:240, 10:250] = 1
        for lt in ("iou", "dice", "tversky", "focal"):
            for ls in (1, 2, 3):
                opt = R.default_options(loss_type=lt, loss_sum=ls, loss_weight=20, loss2_weight=20, loss_weightS=20, loss2_weightS=20)
                m = bare_baseline(net, opt)
                for w, fn in ((5.0, m.compute_topview_loss), (15.0, m.compute_topview_lossB)):
                    x = big.clone().requires_grad_(True)
                    v = fn(x, lab, torch.Tensor([1.0, w]), opt)
                    (g,) = torch.autograd.grad(v, x)
                    key = "%s_s%d_w%d" % (lt, ls, int(w))
                    out[key] = np.float64(v.item())
                    out[key + "_gsum"] = np.float64(g.abs().sum().item())
    np.savez_compressed(os.path.join(GOLD, "kat_bev_variants.npz"), **out)
    for k in sorted(out):
        print(k, float(out[k]))


if __name__ == "__main__":
    main()
