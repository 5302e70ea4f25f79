"""Minimal ``mmcv.Config`` work-alike: ``Config.fromfile("config/cfg_*.py")`` executes the reference's
Python config files unchanged and exposes their top-level names as attributes / items (train.py:51)."""
from __future__ import annotations

import os
import runpy


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


class Config(ConfigDict):
    @staticmethod
    def fromfile(path):
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        ns = runpy.run_path(path)
        cfg = Config({k: _wrap(v) for k, v in ns.items() if not k.startswith("__") and not callable(v) and not hasattr(v, "__file__")})
        cfg["filename"] = path
        return cfg
