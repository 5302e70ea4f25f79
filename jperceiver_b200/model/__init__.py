"""Mirror of the reference's ``mono.model`` package: importing it registers ``Baseline`` in ``MONO``."""
from .registry import MONO  # noqa: F401
from .mono_baseline.net import Baseline  # noqa: F401
