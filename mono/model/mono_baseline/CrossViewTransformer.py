from jperceiver_b200.model.mono_baseline.networks import CrossViewTransformer  # noqa: F401
