"""Golden vectors for the validation metrics, from the REAL reference (authoring container only).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.   Usage:  ``python -m oracle.make_golden_eval``

Imports /root/reference/mono/core/evaluation/pixel_error.py by path (it needs only numpy) and runs its own ``compute_errors``,
``mean_IU`` and ``mean_precision`` on seeded inputs; the per-sample hook body (inline in eval_hooks.py:149-224, which imports
mmcv) is evaluated through ``oracle.eval_port.depth_eval_sample`` around the same cv2 / numpy calls.  Writes
``tests/golden/kat_eval.npz``; the inputs are regenerated from the seeds by ``eval_cases`` (shared with the tests)."""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import eval_port as E  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference/mono/core/evaluation/pixel_error.py"


def depth_case(seed, h, w, gh, gw, density=0.2):
    """Smooth disparity field + sparse LiDAR-like ground truth consistent with it up to a scale and noise."""
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    disp = (0.05 + 0.5 * (yy / h) ** 2 + 0.05 * np.sin(xx / 7.0) * (yy / h) + 0.02 * r.rand(h, w)).astype(np.float32)
    import cv2
    depth = 1.0 / (0.01 + 9.99 * cv2.resize(disp, (gw, gh)))
    gt = (depth * 31.7 * (1 + 0.1 * r.randn(gh, gw))).astype(np.float32)
    gt[r.rand(gh, gw) > density] = 0
    gt[gt > 120] = 120.0            # some values beyond MAX_DEPTH
    return disp, np.abs(gt).astype(np.float32)


def bev_case(seed, occ, kind):
    r = np.random.RandomState(seed)
    true = np.zeros((occ, occ), np.float32)
    pred = np.zeros((occ, occ), np.int64)
    if kind == "mixed":
        true[occ // 4: occ // 2 + 5, occ // 8: occ // 2] = 1
        pred[occ // 4 + 3: occ // 2 + 9, occ // 8 + 2: occ // 2 - 1] = 1
        flip = r.rand(occ, occ) < 0.03
        pred[flip] = 1 - pred[flip]
    elif kind == "pred_empty":
        true[2:9, 3:8] = 1
    elif kind == "gt_empty":
        pred[5:12, 1:6] = 1
    elif kind == "all_one":
        true[:] = 1
        pred[:] = 1
    elif kind == "gt_all_one":
        true[:] = 1
        pred[3:20, :] = 1
    elif kind == "both_empty":
        pass
    return pred, true


DEPTH_CASES = [(0, 24, 80, 47, 155, 0.3), (1, 96, 320, 375, 1242, 0.05), (2, 32, 32, 32, 32, 1.0), (3, 40, 128, 20, 64, 0.5)]
BEV_KINDS = ["mixed", "pred_empty", "gt_empty", "all_one", "gt_all_one", "both_empty"]


def main():
    spec = importlib.util.spec_from_file_location("ref_pixel_error", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    r = np.random.RandomState(7)
    gt = (1 + 60 * r.rand(5000)).astype(np.float32)
    pred = (gt * (1 + 0.2 * r.randn(5000))).clip(1e-3, 80).astype(np.float32)
    mine, theirs = E.compute_errors(gt, pred), ref.compute_errors(gt, pred)
    assert np.allclose(mine, theirs, rtol=0, atol=0), (mine, theirs)
    out["errors_seed7"] = np.array(theirs, np.float64)
    for i, kind in enumerate(BEV_KINDS):
        p, t = bev_case(10 + i, 32, kind)
        with np.errstate(all="ignore"):
            iu, mp = ref.mean_IU(p, t), ref.mean_precision(p, t)
        assert list(map(float, iu)) == list(map(float, E.mean_IU(p, t))), (kind, iu, E.mean_IU(p, t))
        assert list(map(float, mp)) == list(map(float, E.mean_precision(p, t))), (kind, mp, E.mean_precision(p, t))
        out["iu_" + kind] = np.array(iu, np.float64)
        out["mp_" + kind] = np.array(mp, np.float64)
        out["hook_" + kind] = np.array(E.hook_bev_values(p, t), np.float64)
    for c in DEPTH_CASES:
        disp, g = depth_case(*c)
        out["depth_%d" % c[0]] = np.array(E.depth_eval_sample(disp, g), np.float64)
        out["depth_%d_stereo" % c[0]] = np.array(E.depth_eval_sample(disp, g, stereo_scale=True), np.float64)
    np.savez_compressed(os.path.join(GOLD, "kat_eval.npz"), **out)
    for k in sorted(out):
        print(k, out[k])


if __name__ == "__main__":
    main()
