"""Mirror of the reference's ``mono.core`` (mono/core/__init__.py:5-6): the validation hooks with their metrics
and the gradient-exchange names of ``mono.core.utils`` (§8 a-18)."""
from .evaluation import *  # noqa: F401,F403
from .utils import DistOptimizerHook, allreduce_grads  # noqa: F401
