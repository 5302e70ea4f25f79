from jperceiver_b200.core.evaluation import *  # noqa: F401,F403
