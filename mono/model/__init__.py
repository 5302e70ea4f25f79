from jperceiver_b200.model import MONO, Baseline  # noqa: F401
