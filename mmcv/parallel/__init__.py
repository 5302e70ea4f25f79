"""``mmcv.parallel.collate`` as the loader uses it (build_loader.py:8,50)."""
from jperceiver_b200.datasets.loader import collate  # noqa: F401
