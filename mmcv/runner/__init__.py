"""``mmcv.runner`` names used on the training path: ``load_checkpoint`` (train.py:86), ``get_dist_info`` (build_loader.py:27)."""
from collections import OrderedDict

import torch

from jperceiver_b200.apis.env import get_dist_info  # noqa: F401


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    """mmcv 0.4.4 ``load_checkpoint``: a ``{'meta','state_dict','optimizer'}`` file or a bare state_dict, ``module.`` prefix
    stripped; returns the checkpoint."""
    ck = torch.load(filename, map_location=map_location, weights_only=False)
    sd = ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    if sd and all(k.startswith("module.") for k in sd):
        sd = OrderedDict((k[7:], v) for k, v in sd.items())
    target = model.module if hasattr(model, "module") else model
    target.load_state_dict(sd, strict=strict)
    return ck
