from jperceiver_b200.model.mono_baseline.networks import DepthEncoder  # noqa: F401
