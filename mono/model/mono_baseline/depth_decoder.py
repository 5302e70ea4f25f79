from jperceiver_b200.model.mono_baseline.networks import DepthDecoder  # noqa: F401
