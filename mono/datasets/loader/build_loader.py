from jperceiver_b200.datasets.loader import build_dataloader  # noqa: F401
