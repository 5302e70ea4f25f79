"""Import alias for the two ``mmcv`` (0.4.4) names the reference's ``train.py`` uses (``train.py:4-5``): ``Config`` and
``mmcv.runner.load_checkpoint``, plus ``get_dist_info`` / ``collate`` used by its loader.  They resolve to this repository's own
implementations (``jperceiver_b200.apis``); nothing of mmcv is vendored.  With the repository root in front of the reference on
``PYTHONPATH`` the reference's ``train.py`` runs unchanged (``tests/test_train_py_dropin.py``).  If a real mmcv is installed, put
it first on the path instead — the ``mono`` alias package does not need this shim."""
from jperceiver_b200.apis.config import Config, ConfigDict  # noqa: F401

__version__ = "0.4.4+jpb200.shim"
