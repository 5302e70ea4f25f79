from jperceiver_b200.model.mono_baseline.networks import Decoder, Encoder  # noqa: F401
