from jperceiver_b200.core.evaluation.pixel_error import *  # noqa: F401,F403
