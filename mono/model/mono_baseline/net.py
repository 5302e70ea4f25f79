from jperceiver_b200.model.mono_baseline.net import Baseline  # noqa: F401
