from jperceiver_b200.model.mono_baseline.networks import PoseEncoder  # noqa: F401
