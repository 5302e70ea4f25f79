"""Helper of tests/test_train_py_dropin.py (not a test): installs the kernels' host emulation in this process, then executes the
reference's unmodified train.py with the given command line — what ``python train.py ...`` does on a GPU box, minus the GPU."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))     # repository root first: `mono`, `mmcv` -> the alias packages
sys.path.insert(0, HERE)
from emu import build_emulation  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    import torch
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // int(os.environ.get("WORLD_SIZE", "1"))))
    _lib.use_library(build_emulation(), emulated=True)
    train_py = sys.argv[1]
    sys.argv = [train_py] + sys.argv[2:]
    runpy.run_path(train_py, run_name="__main__")
