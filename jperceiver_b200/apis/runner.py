"""Epoch-based runner with the hooks the reference registers.

The reference drives training through mmcv 0.4.4's ``Runner`` (``mono/apis/trainer.py:146-199``):
``register_training_hooks(lr_config, optimizer_config, checkpoint_config, log_config)``, ``resume`` / ``load_checkpoint``,
``run(data_loaders, workflow, total_epochs)``.  mmcv is not part of this path's boundary, so the pieces the configs use are
re-hosted here on top of ``TrainEngine`` with the same file formats:

* learning rate: ``policy='step'`` (``step`` int or list, ``gamma``), optional ``warmup`` in {constant, linear, exp}
  (mmcv ``LrUpdaterHook`` / ``StepLrUpdaterHook`` arithmetic), applied per epoch / per warm-up iteration;
* checkpoints: ``epoch_{n}.pth`` + ``latest.pth`` holding ``{'meta': {'epoch', 'iter', ...}, 'state_dict', 'optimizer'}`` with the
  optimizer in ``torch.optim.Adam.state_dict()`` layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so files
  interchange with the reference's; ``module.`` prefixes of DataParallel checkpoints are accepted on load;
* text log: one JSON object per ``interval`` iterations in ``<work_dir>/<timestamp>.log.json`` (mode, epoch, iter, lr, time,
  data_time, memory, averaged loss terms), the format mmcv's ``TextLoggerHook`` writes.

The optimizer step itself (zero_grad -> backward -> all-reduce -> clip -> Adam, ``DistOptimizerHook.after_train_iter``) is
``TrainEngine.step``; ``optimizer_config['grad_clip']`` is passed to it at construction.
"""
from __future__ import annotations

import json
import logging
import os
import shutil
import time
from collections import OrderedDict

import torch

from .trainer import TrainEngine, change_input_variable


# ------------------------------------------------------------------------------------------------ learning-rate policy
class StepLrPolicy:
    """mmcv ``StepLrUpdaterHook`` (by epoch) with ``LrUpdaterHook`` warm-up."""

    def __init__(self, base_lr, step, gamma=0.1, warmup=None, warmup_iters=0, warmup_ratio=0.1, policy="step", **_unused):
        if policy != "step":
            raise NotImplementedError("lr policy %r: the reference configs use policy='step'" % (policy,))
        if warmup is not None and warmup not in ("constant", "linear", "exp"):
            raise ValueError('"%s" is not a supported type for warming up, valid types are "constant" and "linear"' % warmup)
        if warmup is not None:
            assert warmup_iters > 0 and 0 < warmup_ratio <= 1.0
        if isinstance(step, (list, tuple)):
            assert all(isinstance(s, int) and s > 0 for s in step)
        else:
            assert isinstance(step, int) and step > 0
        self.base_lr, self.step, self.gamma = float(base_lr), step, float(gamma)
        self.warmup, self.warmup_iters, self.warmup_ratio = warmup, int(warmup_iters), float(warmup_ratio)

    def regular_lr(self, epoch):
        if isinstance(self.step, int):
            return self.base_lr * self.gamma ** (epoch // self.step)
        exp = len(self.step)
        for i, s in enumerate(self.step):
            if epoch < s:
                exp = i
                break
        return self.base_lr * self.gamma ** exp

    def lr(self, epoch, cur_iter):
        """Learning rate in effect for iteration ``cur_iter`` (0-based, global) of epoch ``epoch``."""
        reg = self.regular_lr(epoch)
        if self.warmup is None or cur_iter >= self.warmup_iters:
            return reg
        if self.warmup == "constant":
            return reg * self.warmup_ratio
        if self.warmup == "linear":
            k = (1 - cur_iter / self.warmup_iters) * (1 - self.warmup_ratio)
            return reg * (1 - k)
        return reg * self.warmup_ratio ** (1 - cur_iter / self.warmup_iters)


# ------------------------------------------------------------------------------------------------ checkpoint format
def optimizer_state_dict(engine):
    """FusedAdam's flat moments as a ``torch.optim.Adam.state_dict()`` (one param group, parameters in ``model.parameters()``
    order — every parameter is trainable, ``FlatParameters`` checks it — so indices agree with the reference's optimizer).
    Like ``torch.optim.Adam``, parameters that never received a gradient (the three unused ResNet ``fc`` layers, the CCT
    ``res_conv``) have no state entry: their second moment is still exactly zero.  ``initial_lr`` is stored as mmcv's
    ``LrUpdaterHook.before_run`` does, so a resume (here or in the reference) restarts the schedule from the base rate."""
    opt, flat = engine.optimizer, engine.flat
    step = int(opt.step_count.item())
    state = {}
    if step > 0:
        m_host, v_host = opt.exp_avg.detach().cpu(), opt.exp_avg_sq.detach().cpu()      # two copies, then host-side views
        for i, (p, (off, n)) in enumerate(zip(flat.params, flat.views)):
            v = flat._view(v_host, off, p)
            if not bool(v.any()):
                continue
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": flat._view(m_host, off, p).clone().contiguous(),
                        "exp_avg_sq": v.clone().contiguous()}
    group = {"lr": opt.lr, "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": opt.weight_decay, "amsgrad": False,
             "initial_lr": float(getattr(opt, "initial_lr", opt.lr)), "params": list(range(len(flat.params)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(engine, sd):
    opt, flat = engine.optimizer, engine.flat
    groups = sd.get("param_groups", [])
    order = [i for g in groups for i in g["params"]]
    if len(order) != len(flat.params):
        raise ValueError("optimizer state has %d parameters, the model has %d trainable ones" % (len(order), len(flat.params)))
    step = 0
    opt.exp_avg.zero_()
    opt.exp_avg_sq.zero_()                      # parameters without a state entry never received a gradient
    for pos, key in enumerate(order):
        st = sd["state"].get(key)
        if st is None:
            continue
        p, (off, n) = flat.params[pos], flat.views[pos]
        flat._view(opt.exp_avg, off, p).copy_(st["exp_avg"].to(opt.exp_avg.device))
        flat._view(opt.exp_avg_sq, off, p).copy_(st["exp_avg_sq"].to(opt.exp_avg.device))
        step = max(step, int(float(st["step"])))
    opt.step_count.fill_(step)
    if groups:
        opt.lr = float(groups[0].get("lr", opt.lr))
        if "initial_lr" in groups[0]:
            opt.initial_lr = float(groups[0]["initial_lr"])


def save_checkpoint(engine, path, meta):
    model = engine.model
    sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    meta = dict(meta)
    meta.setdefault("time", time.asctime())
    meta.setdefault("writer", "jperceiver_b200")
    torch.save({"meta": meta, "state_dict": sd, "optimizer": optimizer_state_dict(engine)}, path)


def load_checkpoint(engine, path, strict=True, map_location="cpu"):
    ck = torch.load(path, map_location=map_location, weights_only=False)
    sd = ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    if sd and all(k.startswith("module.") for k in sd):
        sd = OrderedDict((k[7:], v) for k, v in sd.items())
    engine.model.load_state_dict(sd, strict=strict)     # parameters are views of the flat buffer: copied in place
    return ck


# ------------------------------------------------------------------------------------------------ runner
class LogBuffer:
    """The two members of ``mmcv.runner.LogBuffer`` the evaluation hooks write (eval_hooks.py:311-327)."""

    def __init__(self):
        self.output = OrderedDict()
        self.ready = False

    def clear_output(self):
        self.output.clear()
        self.ready = False


class Runner:
    def __init__(self, model, optimizer_cfg=None, optimizer_config=None, work_dir=None, log_level="INFO", logger=None, engine=None):
        # a dict (cfg.optimizer_config) or a DistOptimizerHook built from it (trainer.py:183: DistOptimizerHook(**cfg.optimizer_config))
        # ``grad_clip=None`` in the config means NO clipping, as in the reference's DistOptimizerHook (dist_utils.py:56-58); only a
        # runner built without any optimizer_config falls back to the engine default (the configs' max_norm=35)
        if optimizer_config is None:
            self.engine = engine if engine is not None else TrainEngine(model, optimizer_cfg)
        else:
            grad_clip = optimizer_config.grad_clip if hasattr(optimizer_config, "grad_clip") else dict(optimizer_config).get("grad_clip")
            self.engine = engine if engine is not None else TrainEngine(model, optimizer_cfg, grad_clip)
        self.model = self.engine.model
        self.work_dir = work_dir
        if work_dir:
            os.makedirs(work_dir, exist_ok=True)
        self.logger = logger or logging.getLogger("jperceiver_b200")
        self.logger.setLevel(log_level if isinstance(log_level, int) else getattr(logging, str(log_level), logging.INFO))
        self.epoch, self.iter, self.inner_iter = 0, 0, 0
        self.lr_policy = None
        self.checkpoint_interval = None
        self.log_interval = None
        self.json_log = None
        self._buffer = []          # device loss vectors of the iterations since the last log line
        self.log_buffer = LogBuffer()
        self._hooks = []           # extra hooks (``register_hook``): objects with ``after_train_epoch(runner)``
        self._t_iter = self._t_data = 0.0
        self.timestamp = time.strftime("%Y%m%d_%H%M%S", time.localtime())

    # ---- mmcv Runner API used by the reference
    def register_training_hooks(self, lr_config, optimizer_config=None, checkpoint_config=None, log_config=None):
        self.lr_policy = StepLrPolicy(getattr(self.engine.optimizer, "initial_lr", self.engine.optimizer.lr), **dict(lr_config)) if lr_config else None
        if checkpoint_config is not None:
            self.checkpoint_interval = int(dict(checkpoint_config).get("interval", 1))
        if log_config is not None:
            lc = dict(log_config)
            self.log_interval = int(lc.get("interval", 50))
            kinds = [h.get("type") for h in lc.get("hooks", [])]
            unknown = [k for k in kinds if k not in ("TextLoggerHook",)]
            if unknown:
                raise NotImplementedError("logger hooks %r (the reference configs use TextLoggerHook)" % (unknown,))
            if self.work_dir and "TextLoggerHook" in kinds:
                self.json_log = os.path.join(self.work_dir, "%s.log.json" % self.timestamp)

    def register_hook(self, hook, priority="NORMAL"):
        """``mmcv.Runner.register_hook`` for epoch-level hooks (the reference registers its ``DistEvalMonoHook`` this way,
        trainer.py:186-190)."""
        if not hasattr(hook, "after_train_epoch"):
            raise TypeError("hook must define after_train_epoch(runner)")
        self._hooks.append(hook)

    @property
    def rank(self):
        return self.engine.rank

    @property
    def world_size(self):
        return self.engine.world

    def _after_train_epoch(self):
        if self._hooks:
            self.engine.sync_buffers()      # validation scores rank 0's BatchNorm statistics (the ones in the checkpoint) on every rank
        for hook in self._hooks:
            hook.after_train_epoch(self)
        if self.log_buffer.ready:
            rec = OrderedDict(mode="val", epoch=self.epoch + 1, iter=self.inner_iter + 1, lr=self.current_lr()[0])
            rec.update((k, float(v)) for k, v in self.log_buffer.output.items())
            self.logger.info("Epoch(val) [%d][%d]\t%s", rec["epoch"], rec["iter"],
                             ", ".join("%s: %.4f" % (k, float(v)) for k, v in self.log_buffer.output.items()))
            if self.json_log and self.engine.rank == 0:
                with open(self.json_log, "a") as f:
                    f.write(json.dumps(rec) + "\n")
            self.last_val = rec
            self.log_buffer.clear_output()

    def current_lr(self):
        return [self.engine.optimizer.lr]

    def save_checkpoint(self, out_dir=None, filename_tmpl="epoch_{}.pth", meta=None, create_symlink=True):
        out_dir = out_dir or self.work_dir
        m = dict(epoch=self.epoch + 1, iter=self.iter)
        if meta:
            m.update(meta)
        path = os.path.join(out_dir, filename_tmpl.format(self.epoch + 1))
        if self.engine.rank == 0:
            save_checkpoint(self.engine, path, m)
            if create_symlink:
                latest = os.path.join(out_dir, "latest.pth")
                if os.path.lexists(latest):
                    os.remove(latest)
                try:
                    os.symlink(os.path.basename(path), latest)
                except OSError:
                    shutil.copyfile(path, latest)
        return path

    def load_checkpoint(self, filename, strict=False):
        self.logger.info("load checkpoint from %s", filename)
        return load_checkpoint(self.engine, filename, strict=strict)

    def resume(self, checkpoint, resume_optimizer=True):
        ck = self.load_checkpoint(checkpoint, strict=True)
        self.epoch, self.iter = int(ck["meta"]["epoch"]), int(ck["meta"]["iter"])
        if "optimizer" in ck and resume_optimizer:
            load_optimizer_state_dict(self.engine, ck["optimizer"])
            if self.lr_policy is not None:          # the schedule restarts from the BASE rate, not from the decayed one in 'lr'
                self.lr_policy.base_lr = float(self.engine.optimizer.initial_lr)
        self.engine._graph = None
        self.logger.info("resumed epoch %d, iter %d", self.epoch, self.iter)

    # ---- training
    def _set_lr(self):
        if self.lr_policy is not None:
            lr = self.lr_policy.lr(self.epoch, self.iter)
            if lr != self.engine.optimizer.lr:
                self.engine.optimizer.lr = lr
                self.engine._graph = None          # the learning rate is a launch argument of the captured optimizer kernel

    def _log(self, n_iters):
        if not self._buffer:
            return
        vals = torch.stack(self._buffer).mean(0).tolist()          # ONE device->host copy per log line
        self._buffer = []
        rec = OrderedDict(mode="train", epoch=self.epoch + 1, iter=self.inner_iter + 1, lr=self.current_lr()[0])
        if torch.cuda.is_available():
            rec["memory"] = int(torch.cuda.max_memory_allocated() / (1024 * 1024))
        rec["time"] = self._t_iter / max(n_iters, 1)
        rec["data_time"] = self._t_data / max(n_iters, 1)
        for k, v in zip(self.engine.last_names, vals):
            rec[k] = round(float(v), 5)
        self._t_iter = self._t_data = 0.0
        self.logger.info("Epoch [%d][%d]\tlr: %.5g, %s", rec["epoch"], rec["iter"], rec["lr"],
                         ", ".join("%s: %.4f" % (k, rec[k]) for k in self.engine.last_names))
        if self.json_log and self.engine.rank == 0:
            with open(self.json_log, "a") as f:
                f.write(json.dumps(rec) + "\n")
        return rec

    def train_epoch(self, data_loader):
        self.model.train()
        if hasattr(data_loader, "sampler") and hasattr(data_loader.sampler, "set_epoch"):
            data_loader.sampler.set_epoch(self.epoch)              # DistSamplerSeedHook
        since_log = 0
        t0 = time.time()
        for i, batch in enumerate(data_loader):
            self.inner_iter = i
            t1 = time.time()
            self._t_data += t1 - t0
            self._set_lr()
            device = next(self.model.parameters()).device
            data = change_input_variable(batch, device) if device.type == "cuda" else batch
            out = self.engine.step(data, need_log=True)
            self._buffer.append(out.detach())
            self.iter += 1
            since_log += 1
            self._t_iter += time.time() - t0
            if self.log_interval and (i + 1) % self.log_interval == 0:
                self._log(since_log)
                since_log = 0
            t0 = time.time()
        self._buffer = []
        if self.checkpoint_interval and (self.epoch + 1) % self.checkpoint_interval == 0 and self.work_dir:
            self.save_checkpoint(self.work_dir)
        self._after_train_epoch()
        self.epoch += 1

    def run(self, data_loaders, workflow=(("train", 1),), max_epochs=1):
        for mode, _ in workflow:
            if mode != "train":
                raise NotImplementedError("workflow mode %r (the reference configs use [('train', 1)])" % (mode,))
        self.logger.info("Start running, work_dir: %s, max: %d epochs", self.work_dir, max_epochs)
        while self.epoch < max_epochs:
            for (mode, epochs), loader in zip(workflow, data_loaders):
                for _ in range(epochs):
                    if self.epoch >= max_epochs:
                        return
                    self.train_epoch(loader)
