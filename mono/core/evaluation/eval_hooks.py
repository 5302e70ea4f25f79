from jperceiver_b200.core.evaluation.eval_hooks import *  # noqa: F401,F403
