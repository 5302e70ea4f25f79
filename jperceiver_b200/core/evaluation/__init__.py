"""``mono.core.evaluation`` (mono/core/evaluation/__init__.py): metrics and the validation hooks, computed on the device."""
from .pixel_error import AverageMeter, bev_counts, compute_errors, depth_errors, disp_to_depth, mean_IU, mean_precision  # noqa: F401
from .eval_hooks import MAX_DEPTH, MIN_DEPTH, DistEvalHook, DistEvalMonoHook, NonDistEvalHook, eval_crop  # noqa: F401
