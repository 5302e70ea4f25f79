from jperceiver_b200.datasets.get_dataset import SyntheticSnippets, get_dataset  # noqa: F401
