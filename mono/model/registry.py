from jperceiver_b200.model.registry import MONO, Registry  # noqa: F401
