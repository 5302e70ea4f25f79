"""``mono.core.utils.dist_utils`` (dist_utils.py:12-60): the gradient exchange of data-parallel training and the optimizer hook.

The training step of this package does the exchange inside ``TrainEngine.step`` — parameters and gradients live in two flat
buffers, so it is ONE ``all_reduce`` of the flat gradient (NCCL over NVLink5 / NVSwitch), the ``1/world`` scaling rides in the
fused clip + Adam kernel (``apis/trainer.py``).  The two public names of the reference module are kept for code that calls them
directly:

* ``allreduce_grads(model, coalesce=True, bucket_size_mb=-1)`` — same result (every ``param.grad`` replaced by the mean over
  ranks): the gradients are packed once per dtype, reduced with one collective, scaled and scattered back (``TrainEngine`` does
  not need this: its gradients already are one flat buffer).  ``bucket_size_mb`` is accepted and
  ignored: NVSwitch collectives are sized for launch latency, not link count, so one bucket is the right size.
* ``DistOptimizerHook(grad_clip=None, coalesce=True, bucket_size_mb=-1)`` — carries the same options; ``Runner`` reads
  ``grad_clip`` from it.  Its ``after_train_iter`` is what the engine has already done when the hook fires, so it only performs
  the sequence (backward, exchange, clip, step) for a runner that does not own a ``TrainEngine``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_grads(model, coalesce=True, bucket_size_mb=-1):
    world = _world()
    grads = [p.grad.data for p in model.parameters() if p.requires_grad and p.grad is not None]
    if world == 1 or not grads:
        return
    if not coalesce:
        for g in grads:
            dist.all_reduce(g.div_(world))
        return
    by_type = {}
    for g in grads:
        by_type.setdefault((g.dtype, g.device), []).append(g)
    for bucket in by_type.values():
        packed = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(packed)
        packed.div_(world)
        torch._foreach_copy_(bucket, [c.view_as(g) for c, g in zip(packed.split([g.numel() for g in bucket]), bucket)])


class DistOptimizerHook:
    def __init__(self, grad_clip=None, coalesce=True, bucket_size_mb=-1):
        self.grad_clip = grad_clip
        self.coalesce = coalesce
        self.bucket_size_mb = bucket_size_mb

    def clip_grads(self, params):
        return torch.nn.utils.clip_grad_norm_([p for p in params if p.requires_grad and p.grad is not None], **self.grad_clip)

    def after_train_iter(self, runner):
        if getattr(runner, "engine", None) is not None:
            return                              # TrainEngine.step already ran backward + exchange + clip + Adam, fused
        runner.optimizer.zero_grad()
        runner.outputs["loss"].backward()
        allreduce_grads(runner.model, self.coalesce, self.bucket_size_mb)
        if self.grad_clip is not None:
            self.clip_grads(runner.model.parameters())
        runner.optimizer.step()
