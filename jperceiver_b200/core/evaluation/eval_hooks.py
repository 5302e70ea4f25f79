"""Validation hooks (``mono/core/evaluation/eval_hooks.py``) with the metrics computed on the device.

Differences from the reference, none of them in the numbers: no per-sample ``.cpu()`` / cv2 / numpy round trip (the metric
rows stay on the device and are read once per validation pass); ranks exchange their rows with one all-reduce instead of
pickle files in ``work_dir`` (eval_hooks.py:239-257)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

from .pixel_error import AverageMeter, bev_counts, depth_errors, hook_values

MIN_DEPTH = 1e-3
MAX_DEPTH = 80

KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3", "scale", "iou_road", "mAP_road", "iou_vehicle", "mAP_vehicle")


def eval_crop(gt_height, gt_width):
    """eval_hooks.py:168-169 (the Garg / Eigen crop, truncated to int32)."""
    return np.array([0.40810811 * gt_height, 0.99189189 * gt_height, 0.03594771 * gt_width, 0.96405229 * gt_width]).astype(np.int32)


def change_input_variable(data):
    """eval_hooks.py:17-20."""
    for k, v in data.items():
        data[k] = torch.as_tensor(v).float()
    return data


class Hook:
    """The two members of ``mmcv.runner.Hook`` the evaluation hooks use."""

    def every_n_epochs(self, runner, n):
        return (runner.epoch + 1) % n == 0 if n > 0 else False

    def after_train_epoch(self, runner):
        pass


def _device_of(model):
    return next(model.parameters()).device


def evaluate_sample(model, data, device, stereo_scale=False):
    """One validation sample (a dataset item: un-batched tensors) -> float64 device row in ``KEYS`` order
    (eval_hooks.py:140-224)."""
    batch = {k: v.unsqueeze(0).to(device, non_blocking=True) for k, v in data.items()}
    with torch.no_grad():
        result = model(batch)
    row = torch.zeros(len(KEYS), dtype=torch.float64, device=device)
    if "gt_depth" in data:
        row[:8] = depth_errors(result[("disp", 0, 0)], batch["gt_depth"], min_depth=MIN_DEPTH, max_depth=MAX_DEPTH,
                               stereo_scale=stereo_scale)[0]
    for j, (out_key, lab_key) in enumerate((("topview", ("bothS", 0, 0)), ("topviewB", ("bothD", 0, 0)))):
        if out_key in result and lab_key in batch:
            lg = result[out_key]
            iou, mAP = hook_values(bev_counts(lg, batch[lab_key]), lg.shape[-1] * lg.shape[-2])
            row[8 + 2 * j], row[9 + 2 * j] = iou[0], mAP[0]
    return row


class DistEvalHook(Hook):
    """eval_hooks.py:98-262: every ``interval`` epochs run the model in eval mode over ``dataset`` (sample ``idx`` on rank
    ``idx % world_size``), collect one metric row per sample and hand the list of per-sample dicts to ``evaluate``."""

    def __init__(self, dataset, interval=1, cfg=None):
        if not (hasattr(dataset, "__len__") and hasattr(dataset, "__getitem__")):
            raise TypeError("dataset must be a map-style Dataset")
        self.dataset = dataset
        self.interval = interval
        self.cfg = cfg
        self.count = 0

    def _stereo_scale(self):
        data_cfg = (self.cfg or {}).get("data", {}) if hasattr(self.cfg, "get") else {}
        return bool(data_cfg.get("stereo_scale", False))

    def after_train_epoch(self, runner):
        self.count += 1
        if not self.every_n_epochs(runner, self.interval):
            return
        runner.model.eval()
        rank = getattr(runner, "rank", 0)
        world = getattr(runner, "world_size", 1)
        device = _device_of(runner.model)
        n = len(self.dataset)
        rows = torch.zeros(n, len(KEYS), dtype=torch.float64, device=device)
        for idx in range(rank, n, world):
            rows[idx] = evaluate_sample(runner.model, change_input_variable(self.dataset[idx]), device, self._stereo_scale())
        if world > 1 and dist.is_available() and dist.is_initialized():
            dist.all_reduce(rows)                                    # every row is written by exactly one rank
        runner.model.train()
        if rank == 0:
            table = rows.tolist()                                    # the ONE device->host copy of the pass
            self.evaluate(runner, [dict(zip(KEYS, r)) for r in table])

    def evaluate(self, runner, results):
        raise NotImplementedError


class DistEvalMonoHook(DistEvalHook):
    """eval_hooks.py:265-327: average every metric over the samples into ``runner.log_buffer.output``."""

    def evaluate(self, runner, results):
        if not isinstance(results, list):
            raise TypeError("results must be a list of per-sample dicts, not {}".format(type(results)))
        meters = OrderedDict((k, AverageMeter()) for k in KEYS)
        for result in results:
            for k in KEYS:
                meters[k].update(result[k])
        for k in KEYS:
            runner.log_buffer.output["scale mean" if k == "scale" else k] = meters[k].avg
        runner.log_buffer.ready = True


class NonDistEvalHook(Hook):
    """eval_hooks.py:27-95: single-process depth-only evaluation; prints the running and final ``a1``."""

    def __init__(self, dataset, cfg):
        self.dataset = dataset
        self.interval = cfg.get("interval", 1)
        self.out_path = cfg.get("work_dir", "./")
        self.cfg = cfg

    def after_train_epoch(self, runner):
        if not self.every_n_epochs(runner, self.interval):
            return
        runner.model.eval()
        device = _device_of(runner.model)
        rows = [evaluate_sample(runner.model, change_input_variable(self.dataset[idx]), device) for idx in range(len(self.dataset))]
        runner.model.train()
        table = torch.stack(rows).tolist() if rows else []
        meters = OrderedDict((k, AverageMeter()) for k in KEYS[:7])
        for r in table:
            for k, v in zip(KEYS[:7], r):
                meters[k].update(v)
            print("a1_ is ", r[4])
        print("a1 is ", meters["a1"].avg)
        self.results = OrderedDict((k, m.avg) for k, m in meters.items())
