from jperceiver_b200.model.mono_baseline.networks import CycledViewProjection  # noqa: F401
