from jperceiver_b200.model.mono_baseline.networks import ResnetEncoder  # noqa: F401
