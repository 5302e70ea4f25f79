"""Import the reference's own hot-path modules from ``/root/reference`` (authoring container only).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Nothing is copied: the reference's files are
executed where they lie, under synthetic package names so they cannot collide with this repo's
``mono`` alias package:

* ``_jref``       -> ``/root/reference/mono``  with ``mono/model/mono_baseline/net.py``
                    (the registered copy: only ``type="Argo_both"`` runs)
* ``_jref_root``  -> same tree, but ``net`` is loaded from ``/root/reference/net.py``
                    (the alternate copy: ``static`` / ``static_raw`` / ``dynamic`` / ``Argo_*``)

Shims, all oracle-side:
  - stub modules ``matplotlib(.pyplot/.cm)``, ``imageio``, ``pykitti`` (imported, unused on the path;
    ``plt.figure()`` is called at net.py:221,414 and must exist);
  - ``skimage.segmentation.find_boundaries`` restated (mode='inner', connectivity=1) — parity unpinned;
  - ``torchgeometry.core.{imgwarp.warp_perspective, transformations.transform_points}`` restated from
    the published 0.1.2 algorithm — parity unpinned; ``align_corners`` is an explicit knob
    (``set_warp_align_corners``) because 0.1.2 leaves it to torch's default;
  - ``.cuda()`` on tensors/modules -> identity (the reference hard-codes ``.cuda()``);
  - ``torchvision.models.resnet18(pretrained=True)`` -> offline random init (would download).
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("JPB200_REFERENCE_ROOT", "/root/reference")

_WARP_ALIGN_CORNERS = True


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mono", "model", "mono_baseline", "net.py"))


def set_warp_align_corners(flag: bool) -> None:
    """Pin the un-pinned torchgeometry sampling convention."""
    global _WARP_ALIGN_CORNERS
    _WARP_ALIGN_CORNERS = bool(flag)


# ----------------------------------------------------------------------------------------------
# third-party restatements
# ----------------------------------------------------------------------------------------------
def _tg_transform_points(trans_01: torch.Tensor, points_1: torch.Tensor) -> torch.Tensor:
    """torchgeometry 0.1.2 ``transform_points``: homogeneous multiply then divide by last coord."""
    ones = torch.ones_like(points_1[..., :1])
    ph = torch.cat([points_1, ones], dim=-1)
    out = torch.matmul(trans_01.unsqueeze(1), ph.unsqueeze(-1)).squeeze(-1)
    return out[..., :-1] / out[..., -1:]


def _tg_normal_transform_pixel(h: int, w: int) -> torch.Tensor:
    return torch.tensor([[2.0 / (w - 1), 0.0, -1.0], [0.0, 2.0 / (h - 1), -1.0], [0.0, 0.0, 1.0]])


def _tg_warp_perspective(src, M, dsize, flags="bilinear", border_mode=None, border_value=0):
    """torchgeometry 0.1.2 ``warp_perspective`` (dst_pix <- src_pix homography ``M``)."""
    b, _, hs, ws = src.shape
    hd, wd = dsize
    src_norm_T_src_pix = _tg_normal_transform_pixel(hs, ws).to(M)
    dst_norm_T_dst_pix = _tg_normal_transform_pixel(hd, wd).to(M)
    dst_norm_T_src_norm = dst_norm_T_dst_pix @ (M @ torch.inverse(src_norm_T_src_pix))
    src_norm_T_dst_norm = torch.inverse(dst_norm_T_src_norm)
    ys = torch.linspace(-1, 1, hd)
    xs = torch.linspace(-1, 1, wd)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    grid = torch.stack([gx, gy], dim=-1).view(1, -1, 2).expand(b, -1, -1).to(M)
    flow = _tg_transform_points(src_norm_T_dst_norm, grid).view(b, hd, wd, 2)
    return F.grid_sample(src, flow, mode="bilinear", padding_mode="zeros", align_corners=_WARP_ALIGN_CORNERS)


def _find_boundaries(label_img, connectivity=1, mode="thick", background=0):
    """skimage ``find_boundaries`` for mode='inner', connectivity=1 (the only use, boundary_loss.py:142)."""
    from scipy import ndimage as ndi

    if mode != "inner":
        raise NotImplementedError(mode)
    lab = np.asarray(label_img)
    if lab.dtype == bool:
        lab = lab.astype(np.uint8)
    fp = ndi.generate_binary_structure(lab.ndim, connectivity)
    thick = ndi.grey_dilation(lab, footprint=fp) != ndi.grey_erosion(lab, footprint=fp)
    return thick & (lab != background)


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot", figure=lambda *a, **k: None, imsave=lambda *a, **k: None)
        cm = _stub("matplotlib.cm")
        _stub("matplotlib", pyplot=plt, cm=cm)
    for name in ("imageio", "pykitti"):
        if name not in sys.modules:
            _stub(name)
    if "skimage" not in sys.modules:
        seg = _stub("skimage.segmentation", find_boundaries=_find_boundaries)
        _stub("skimage", segmentation=seg)
    if "torchgeometry" not in sys.modules:
        iw = _stub("torchgeometry.core.imgwarp", warp_perspective=_tg_warp_perspective)
        tr = _stub("torchgeometry.core.transformations", transform_points=_tg_transform_points)
        core = _stub("torchgeometry.core", imgwarp=iw, transformations=tr)
        _stub("torchgeometry", core=core)
    if not hasattr(np, "bool"):
        np.bool = bool  # boundary_loss.py:139 uses the removed alias


@contextlib.contextmanager
def cpu_cuda_identity():
    """Make the reference's hard-coded ``.cuda()`` calls no-ops while it runs on the CPU."""
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def _pkg(name: str, path: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load_file(modname: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE: dict = {}


def load(variant: str = "registered") -> types.ModuleType:
    """Return the reference's ``net`` module.  ``variant``: 'registered' (Argo_both) or 'root' (static*)."""
    if variant in _CACHE:
        return _CACHE[variant]
    if not available():
        raise RuntimeError("reference tree not present at %s (GPU box?)" % REFERENCE_ROOT)
    _install_stubs()
    import torchvision.models as tvm

    if not getattr(tvm.resnet18, "_jref_offline", False):
        orig = tvm.resnet18

        def offline_resnet18(pretrained=False, *a, **k):
            return orig(weights=None)

        offline_resnet18._jref_offline = True
        tvm.resnet18 = offline_resnet18

    top = {"registered": "_jref", "root": "_jref_root"}[variant]
    base = os.path.join(REFERENCE_ROOT, "mono")
    _pkg(top, base)
    _pkg(top + ".model", os.path.join(base, "model"))
    _pkg(top + ".model.mono_baseline", os.path.join(base, "model", "mono_baseline"))
    netfile = (os.path.join(base, "model", "mono_baseline", "net.py") if variant == "registered"
               else os.path.join(REFERENCE_ROOT, "net.py"))
    mod = _load_file(top + ".model.mono_baseline.net", netfile)
    _CACHE[variant] = mod
    return mod


class Options(dict):
    """Attribute + item dict, as mmcv's ConfigDict behaves for ``cfg.model``."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_options(**over) -> Options:
    """Model options of config/cfg_kitti_baseline_argo_both_boundary_ce_iou_1024_20_B1.py:22-54."""
    o = Options(
        name="Baseline", depth_num_layers=18, pose_num_layers=18, frame_ids=[0, -1, 1], imgs_per_gpu=1,
        height=1024, width=1024, scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0,
        depth_pretrained_path=None, pose_pretrained_path=None, automask=True, disp_norm=True,
        smoothness_weight=1e-3, scale_weight=0.1, dynamic_weight=15.0, static_weight=5.0,
        occ_map_size=256, num_class=2, loss_type="iou", loss_weight=20, loss_weightS=20,
        loss2_type="boundary", loss2_weight=20, loss2_weightS=20, type="Argo_both", loss_sum=3, split="argo",
    )
    o.update(over)
    return o


def build_baseline(opt, variant: str | None = None):
    if variant is None:
        variant = "registered" if opt["type"] == "Argo_both" else "root"
    net = load(variant)
    with cpu_cuda_identity():
        model = net.Baseline(opt)
    return model
