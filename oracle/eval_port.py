"""CPU restatement of the reference's validation metrics — TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows /root/reference/mono/core/evaluation/eval_hooks.py:149-224 (the per-sample body of ``DistEvalHook.after_train_epoch``)
and pixel_error.py:27-40, 62-118.  numpy + OpenCV on the host, exactly the libraries the reference calls.

Parity status: ``compute_errors`` / ``mean_IU`` / ``mean_precision`` are pinned against the reference's own functions
(``oracle/make_golden_eval.py`` imports ``pixel_error.py`` from /root/reference and writes ``tests/golden/kat_eval.npz``).  The
hook body is inline code in the reference (not callable on its own; the module imports mmcv, absent here), so
``depth_eval_sample`` restates it line by line around the same ``cv2.resize`` / ``np.median`` calls — its third-party arithmetic
(OpenCV ``INTER_LINEAR``, unpinned ``opencv_python``) is "parity unpinned" in the sense of SURVEY.md §8c."""
from __future__ import annotations

import numpy as np

MIN_DEPTH = 1e-3
MAX_DEPTH = 80


def compute_errors(gt, pred):
    """pixel_error.py:27-40."""
    ratio = np.maximum(gt / pred, pred / gt)
    a = [(ratio < 1.25 ** k).mean() for k in (1, 2, 3)]
    sq = (gt - pred) ** 2
    rmse = np.sqrt(sq.mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    return np.mean(np.abs(gt - pred) / gt), np.mean(sq / gt), rmse, rmse_log, a[0], a[1], a[2]


def depth_eval_sample(disp, gt_depth, stereo_scale=False, min_depth=0.1, max_depth=100.0):
    """eval_hooks.py:158-190 for one sample: ``disp`` h×w float32 (network output), ``gt_depth`` gh×gw.  Returns the 7 errors
    and the median ratio (8 floats)."""
    import cv2
    pred_disp = (1.0 / max_depth + (1.0 / min_depth - 1.0 / max_depth) * disp.astype(np.float32)).astype(np.float32)
    gh, gw = gt_depth.shape[:2]
    pred_depth = 1 / cv2.resize(pred_disp, (gw, gh))
    rows = np.array([0.40810811 * gh, 0.99189189 * gh]).astype(np.int32)
    cols = np.array([0.03594771 * gw, 0.96405229 * gw]).astype(np.int32)
    keep = np.zeros(gt_depth.shape, bool)
    keep[rows[0]:rows[1], cols[0]:cols[1]] = True
    keep &= (gt_depth > MIN_DEPTH) & (gt_depth < MAX_DEPTH)
    p, g = pred_depth[keep], gt_depth[keep]
    ratio = np.median(g) / np.median(p)
    p = p * (36 if stereo_scale else ratio)
    p = np.clip(p, MIN_DEPTH, MAX_DEPTH)
    return tuple(float(v) for v in compute_errors(g, p)) + (float(ratio),)


def _masks(segm, classes):
    return [segm == c for c in classes]


def mean_IU(eval_segm, gt_segm):
    """pixel_error.py:80-118."""
    classes = np.union1d(np.unique(eval_segm), np.unique(gt_segm))
    out = []
    for e, g in zip(_masks(eval_segm, classes), _masks(gt_segm, classes)):
        if e.sum() == 0 or g.sum() == 0:
            out.append(0)
            continue
        inter = np.logical_and(e, g).sum()
        out.append(inter / (g.sum() + e.sum() - inter))
    return out


def mean_precision(eval_segm, gt_segm):
    """pixel_error.py:62-77 (0/0 -> 0)."""
    classes = np.unique(gt_segm)
    out = []
    for e, g in zip(_masks(eval_segm, classes), _masks(gt_segm, classes)):
        n = e.sum()
        out.append(0. if n == 0 else np.logical_and(e, g).sum() / float(n))
    return out


def hook_bev_values(pred, true):
    """eval_hooks.py:185-224: ``acc = np.array([0., 0.]); acc += metric(pred, true); value = acc[1]`` for IoU and precision."""
    iou = np.array([0., 0.])
    iou += mean_IU(pred, true)
    prec = np.array([0., 0.])
    prec += mean_precision(pred, true)
    return float(iou[1]), float(prec[1])
