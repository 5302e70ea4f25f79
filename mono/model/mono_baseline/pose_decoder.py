from jperceiver_b200.model.mono_baseline.networks import PoseDecoder  # noqa: F401
