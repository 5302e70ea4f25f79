"""Device versions of the reference's evaluation metrics (``mono/core/evaluation/pixel_error.py``).

Same names, argument meaning and return values as the reference functions, but the inputs are CUDA tensors and the
arithmetic runs in ``libjpb200.so`` (``csrc/eval.cu``): no ``.cpu()`` of maps, no ``np.unique`` / mask stacks.  There is no
CPU fallback (``_lib.ptr`` refuses CPU tensors)."""
from __future__ import annotations

import ctypes as C

import torch

from ..._lib import DepthEvalArgs, check, lib, ptr, stream_of


class AverageMeter(object):
    """pixel_error.py:7-24."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def disp_to_depth(disp, min_depth=0.1, max_depth=100):
    """pixel_error.py:43-48 — returns ``(scaled_disp, depth)``."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    return scaled_disp, 1 / scaled_disp


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def depth_errors(disp, gt_depth, *, min_depth=1e-3, max_depth=80.0, crop=None, stereo_scale=False, net_min_depth=0.1,
                 net_max_depth=100.0):
    """The per-sample depth evaluation of ``DistEvalHook.after_train_epoch`` (eval_hooks.py:149-197) for a batch, in one call.

    ``disp``: B×1×h×w (or B×h×w) network output ``("disp", 0, 0)``; ``gt_depth``: B×gh×gw.  Returns a float64 tensor B×8 on
    the device: ``abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, ratio`` (NaN rows where no ground-truth pixel is valid, as
    numpy's mean of an empty selection).  ``crop``: ``(r0, r1, c0, c1)``; default = the reference's crop of the gt frame."""
    from .eval_hooks import eval_crop
    d = _f32c(disp)
    if d.dim() == 4:
        d = d[:, 0].contiguous()
    g = _f32c(gt_depth)
    if g.dim() == 2:
        g = g[None]
    B, h, w = d.shape
    if g.shape[0] != B:
        raise ValueError("disp has %d samples, gt_depth %d" % (B, g.shape[0]))
    gh, gw = g.shape[-2:]
    a = DepthEvalArgs()
    a.disp, a.gt = ptr(d), ptr(g)
    a.B, a.h, a.w, a.gh, a.gw = B, h, w, gh, gw
    a.min_disp, a.max_disp = 1.0 / net_max_depth, 1.0 / net_min_depth
    a.min_depth, a.max_depth = float(min_depth), float(max_depth)
    c = eval_crop(gh, gw) if crop is None else crop
    for i in range(4):
        a.crop[i] = int(c[i])
    a.fixed_scale = 36.0 if stereo_scale is True else float(stereo_scale or 0.0)
    work = torch.empty(B, 2, gh * gw, dtype=torch.float32, device=d.device)
    count = torch.zeros(B, dtype=torch.int32, device=d.device)
    out = torch.empty(B, 8, dtype=torch.float64, device=d.device)
    a.work, a.count, a.out = ptr(work), ptr(count), ptr(out)
    check(lib().jpb_depth_eval(C.byref(a), stream_of(d)), "jpb_depth_eval")
    return out


def compute_errors(gt, pred):
    """pixel_error.py:27-40 on 1-D device tensors of already selected, already scaled depths: returns
    ``abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3`` as Python floats (one device->host copy of 7 numbers)."""
    gt, pred = _f32c(gt).reshape(1, 1, -1), _f32c(pred).reshape(1, 1, -1)
    n = gt.shape[-1]
    # the same kernel with an identity "resize" (h,w == gh,gw), an all-pass crop, the depths passed as 1/disp and scale 1
    out = depth_errors(1.0 / pred, gt, min_depth=0.0, max_depth=float("inf"), crop=(0, 1, 0, n), stereo_scale=1.0,
                       net_min_depth=1.0, net_max_depth=float("inf"))
    return tuple(out[0, :7].tolist())


def bev_counts(logits, label):
    """int64 B×3 device tensor ``[#(pred==1 & gt==1), #(pred==1), #(gt==1)]`` with ``pred = argmax(logits, 1)``
    (eval_hooks.py:185-197).  ``logits``: B×2×occ×occ in any dense layout; ``label``: B×1×occ×occ or B×occ×occ."""
    lg = logits.detach()
    if lg.dtype != torch.float32:
        lg = lg.float()
    B, nc, H, W = lg.shape
    if nc != 2 or H != W:
        raise ValueError("logits must be Bx2xoccxocc")
    if not (lg.is_contiguous() or lg.is_contiguous(memory_format=torch.channels_last)):
        lg = lg.contiguous()
    lab = _f32c(label).reshape(B, H * W)
    counts = torch.zeros(B, 3, dtype=torch.int64, device=lg.device)
    sb, sc, sh, sw = lg.stride()
    if sh != W * sw:
        raise ValueError("unsupported logits layout")
    check(lib().jpb_bev_confusion(ptr(lg), sb, sc, sw, ptr(lab), B, H, ptr(counts), stream_of(lg)), "jpb_bev_confusion")
    return counts


def bev_metrics_from_counts(counts, npix):
    """Per-sample IoU / precision of both classes from :func:`bev_counts` rows, on the device (float64 tensors).

    Returns ``(IU[n,2], P[n,2], present[n,2], gt_present[n,2])`` with the reference's conventions: IoU of a class is 0 when the
    class is missing from the prediction or from the ground truth (pixel_error.py:103-104), precision is 0 when the
    prediction never emits it (pixel_error.py:72-75); ``present`` = class occurs in either map (``union_classes``),
    ``gt_present`` = class occurs in the ground truth (``extract_classes(gt_segm)``)."""
    c = counts.to(torch.float64)
    n11, p1, g1 = c[:, 0], c[:, 1], c[:, 2]
    n_ii = torch.stack([npix - p1 - g1 + n11, n11], 1)
    n_ij = torch.stack([npix - p1, p1], 1)
    t_i = torch.stack([npix - g1, g1], 1)
    zero = torch.zeros_like(n_ii)
    bad = (n_ij == 0) | (t_i == 0)
    IU = torch.where(bad, zero, n_ii / torch.where(bad, torch.ones_like(n_ii), t_i + n_ij - n_ii))
    P = torch.where(n_ij == 0, zero, n_ii / torch.where(n_ij == 0, torch.ones_like(n_ii), n_ij))
    return IU, P, (n_ij > 0) | (t_i > 0), t_i > 0


def hook_values(counts, npix):
    """What ``DistEvalHook`` stores per sample as ``iou_*`` / ``mAP_*`` (eval_hooks.py:185-224): ``np.array([0., 0.]) +=
    mean_IU(...)`` then element ``[1]`` — class 1 when two classes are listed, and (numpy broadcasting of a 1-element list)
    the only listed class otherwise.  Returns two float64 device tensors of length n."""
    IU, P, present, gt_present = bev_metrics_from_counts(counts, npix)
    iou = torch.where(present[:, 0] & present[:, 1], IU[:, 1], torch.where(present[:, 0], IU[:, 0], IU[:, 1]))
    mAP = torch.where(gt_present[:, 0] & gt_present[:, 1], P[:, 1], torch.where(gt_present[:, 0], P[:, 0], P[:, 1]))
    return iou, mAP


def _class_counts(eval_segm, gt_segm):
    """:func:`bev_counts` row of two H×W class maps with classes {0, 1} (device tensors)."""
    e, g = eval_segm.detach(), gt_segm.detach()
    if e.shape != g.shape or e.dim() != 2:
        raise ValueError("DiffDim: Different dimensions of matrices!")     # pixel_error.py:175-180 (EvalSegErr)
    if e.shape[0] != e.shape[1]:
        raise ValueError("BEV maps are square (occ x occ)")
    lg = torch.stack([torch.zeros_like(e, dtype=torch.float32), (e != 0).float()])[None]   # two "logits" whose argmax is the map
    return bev_counts(lg, (g != 0).float()[None]), e.numel()


def mean_IU(eval_segm, gt_segm):
    """pixel_error.py:80-118: list of per-class IoU over the classes present in either map (sorted); a class missing from
    the prediction or from the ground truth contributes 0."""
    IU, _, present, _ = bev_metrics_from_counts(*_class_counts(eval_segm, gt_segm))
    IU, present = IU[0].tolist(), present[0].tolist()
    return [IU[c] for c in (0, 1) if present[c]]


def mean_precision(eval_segm, gt_segm):
    """pixel_error.py:62-77: list of per-class precision over the classes present in the GROUND TRUTH; 0 where the
    prediction never emits the class."""
    _, P, _, gt_present = bev_metrics_from_counts(*_class_counts(eval_segm, gt_segm))
    P, gt_present = P[0].tolist(), gt_present[0].tolist()
    return [P[c] for c in (0, 1) if gt_present[c]]
