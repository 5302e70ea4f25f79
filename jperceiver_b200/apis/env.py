"""Process-group / seeding / logging helpers with the reference's names (mono/apis/env.py:17-77).
One process per GPU; NCCL over NVLink/NVSwitch through ``torch.distributed``."""
from __future__ import annotations

import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_dist(launcher="pytorch", backend="nccl", **kwargs):
    if launcher != "pytorch":
        raise ValueError("Invalid launcher type: {} (only the torch.distributed launcher is supported)".format(launcher))
    rank = int(os.environ["RANK"])
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank % max(torch.cuda.device_count(), 1))))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend, **kwargs)


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_root_logger(log_level=logging.INFO):
    logger = logging.getLogger()
    if not logger.hasHandlers():
        logging.basicConfig(format="%(asctime)s - %(levelname)s - %(message)s", level=log_level)
    rank, _ = get_dist_info()
    if rank != 0:
        logger.setLevel("ERROR")
    return logger
