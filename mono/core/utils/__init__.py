from jperceiver_b200.core.utils import DistOptimizerHook, allreduce_grads  # noqa: F401
