"""Synthetic batches with the reference dataset's dict keys and shapes: there are no
KITTI/Argoverse files in this environment.  Frames are a smooth low-frequency field plus noise, shifted
between frame ids, so the warped / identity arg-min branches and the SSIM windows are all exercised."""
from __future__ import annotations

import torch
import torch.nn.functional as F

KITTI_TR = [[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
            [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
            [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01],
            [0.0, 0.0, 0.0, 1.0]]


def make_batch(opt, B, seed=1024, split=None, pin=False):
    """Host (CPU) batch dict: ``("color"|"color_aug", f, 0)`` B×3×H×W, ``("color",0,-1)`` full-res frame (shape
    only), ``("K",0)``/``("inv_K",0)`` B×4×4, ``("odometry_K",0,0)``, ``("Tr_cam2_velo",0,0)``, BEV labels."""
    split = split or opt["split"]
    g = torch.Generator().manual_seed(seed)
    H, W, occ = opt["height"], opt["width"], opt["occ_map_size"]
    hw_full = (2056, 2464) if split == "argo" else (375, 1242)
    d = {}
    base = F.interpolate(torch.rand(B, 3, H // 16 + 2, W // 16 + 2, generator=g), (H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    for f in opt["frame_ids"]:
        img = (0.85 * torch.roll(base, shifts=(2 * f, 5 * f), dims=(2, 3)) + 0.15 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
        d[("color", f, 0)] = img
        d[("color_aug", f, 0)] = img.clone()
    d[("color", 0, -1)] = torch.zeros(B, 3, *hw_full)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    d[("K", 0)] = K.repeat(B, 1, 1)
    d[("inv_K", 0)] = torch.linalg.pinv(K).repeat(B, 1, 1)
    if split == "argo":
        oK = torch.tensor([[1400.0, 0, 1232.0, 0], [0, 1400.0, 1028.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    else:
        oK = torch.tensor([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1.0]])
    d[("odometry_K", 0, 0)] = oK.repeat(B, 1, 1)
    d[("Tr_cam2_velo", 0, 0)] = torch.tensor(KITTI_TR).repeat(B, 1, 1)
    yy, xx = torch.meshgrid(torch.arange(occ), torch.arange(occ), indexing="ij")
    for name, frac in (("bothS", 0.55), ("bothD", 0.12), ("both_dynamic", 0.45)):
        m = torch.zeros(B, 1, occ, occ)
        for b in range(B):
            cy, cx = (torch.rand(2, generator=g) * 0.3 + 0.35) * occ
            ry, rx = (torch.rand(2, generator=g) * 0.5 + 0.5) * frac * occ
            m[b, 0] = ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1).float()
            if name != "bothD":   # the road always covers the strip ahead of the ego car (so the CGT "assumption region" is labelled)
                m[b, 0, int(0.5 * occ):, int(0.3 * occ):int(0.7 * occ)] = 1   # ego car sits at the bottom centre of the BEV map
        d[(name, 0, 0)] = m
    if pin:
        d = {k: v.pin_memory() for k, v in d.items()}
    return d


def batch_bytes(batch):
    return sum(v.numel() * 4 for v in batch.values() if torch.is_tensor(v))
