from jperceiver_b200.core.utils.dist_utils import DistOptimizerHook, allreduce_grads  # noqa: F401
