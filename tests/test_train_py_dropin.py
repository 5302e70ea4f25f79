"""The drop-in boundary, literally: the REFERENCE's own ``train.py`` (unmodified, executed from /root/reference) drives this
repository — its ``mmcv`` / ``mono.*`` imports resolve to the alias packages at the repository root, the config is a file in
the reference's format, the model / loader / runner / checkpoints are this repository's.  Runs on the kernels' host emulation
(no GPU here) at a reduced shape with synthetic snippets; skipped where /root/reference is absent (the GPU box)."""
import json
import os
import runpy
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402

TRAIN_PY = "/root/reference/train.py"
REF_CFG = "/root/reference/config/cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20.py"

pytestmark = pytest.mark.skipif(not os.path.isfile(TRAIN_PY), reason="reference tree not present")


@pytest.fixture()
def emu():
    _lib._handle, _lib._emulated = None, False
    _lib.use_library(build_emulation(), emulated=True)
    yield
    _lib._handle, _lib._emulated = None, False


def _small_config(tmp_path, extra=""):
    """The reference's odometry config, executed as it is, with the shapes reduced and the authors' file paths removed."""
    text = open(REF_CFG).read()
    text += '''
# ---- test overrides (appended; everything above is the reference's file, byte for byte)
HEIGHT, WIDTH, IMGS_PER_GPU = 128, 384, 2
data.update(name='synthetic', height=HEIGHT, width=WIDTH, num_samples=4, occ_map_size=64, frame_ids=[0, -1])
model.update(height=HEIGHT, width=WIDTH, imgs_per_gpu=IMGS_PER_GPU, occ_map_size=64, frame_ids=[0, -1],
             depth_pretrained_path=None, pose_pretrained_path=None)
imgs_per_gpu, workers_per_gpu, total_epochs = IMGS_PER_GPU, 0, 1          # validate = True stays, as in the reference's file
log_config = dict(interval=1, hooks=[dict(type='TextLoggerHook')])
''' + extra
    path = tmp_path / "cfg_small.py"
    path.write_text(text)
    return str(path)


def test_reference_train_py_runs_unchanged(emu, tmp_path, monkeypatch):
    work = str(tmp_path / "work")
    monkeypatch.setattr(sys, "argv", [TRAIN_PY, "--config", _small_config(tmp_path), "--work_dir", work, "--launcher", "none", "--gpus", "0"])
    monkeypatch.syspath_prepend(ROOT)               # `mono` and `mmcv` -> the alias packages of this repository
    for name in [m for m in sys.modules if m == "mmcv" or m.startswith("mmcv.")]:
        monkeypatch.delitem(sys.modules, name)
    limit = sys.getrecursionlimit()
    try:
        runpy.run_path(TRAIN_PY, run_name="__main__")
    finally:
        sys.setrecursionlimit(limit)
    import mmcv
    import mono.apis
    assert os.path.dirname(os.path.abspath(mmcv.__file__)).startswith(ROOT) and mono.apis.__file__.startswith(ROOT)
    # one epoch of 4 snippets at 2 per step: two iterations, a checkpoint in the reference's format, two JSON log lines
    ck = torch.load(os.path.join(work, "epoch_1.pth"), weights_only=False)
    assert set(ck) == {"meta", "state_dict", "optimizer"} and ck["meta"]["epoch"] == 1 and ck["meta"]["iter"] == 2
    assert os.path.exists(os.path.join(work, "latest.pth"))
    logs = [f for f in os.listdir(work) if f.endswith(".log.json")]
    lines = [json.loads(l) for l in open(os.path.join(work, logs[0]))]
    train = [l for l in lines if l["mode"] == "train"]
    assert len(train) == 2 and train[0]["lr"] == 1e-4
    # cfg.validate = True (the reference's setting): DistEvalMonoHook ran over the validation split after the epoch
    val = [l for l in lines if l["mode"] == "val"]
    assert len(val) == 1 and all(k in val[0] for k in ("abs_rel", "a1", "scale mean", "iou_road", "mAP_road")), lines
    assert 0.0 <= val[0]["a1"] <= 1.0 and val[0]["abs_rel"] > 0.0
    for key in ("topview_loss", "transform_topview_loss", "transform_loss", "layout_loss", "loss",
                "('min_reconstruct_loss', 0)", "('scale_loss', 3)", "('smooth_loss', 2)"):
        assert key in lines[0] and lines[0][key] == lines[0][key], key     # present and not NaN
    keys = list(ck["state_dict"])
    assert len(keys) == 766 and keys[0].startswith("DepthEncoder.")


def test_reference_train_py_two_ranks_gloo(tmp_path):
    """``--launcher pytorch`` under torchrun, 2 processes (gloo instead of nccl in ``dist_params``, no GPU here): init_dist, the
    DistributedGroupSampler plan (8 snippets -> 4 per rank -> 2 steps), the gradient all-reduce, rank-0 checkpoint and log."""
    import socket
    import subprocess
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    cfg = _small_config(tmp_path, "data.update(num_samples=8)\ndist_params = dict(backend='gloo')\n")
    work = str(tmp_path / "work2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_run_ref_train.py"), TRAIN_PY, "--config", cfg, "--work_dir", work,
           "--launcher", "pytorch"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ck = torch.load(os.path.join(work, "epoch_1.pth"), weights_only=False)
    assert ck["meta"]["epoch"] == 1 and ck["meta"]["iter"] == 2
    logs = [f for f in os.listdir(work) if f.endswith(".log.json")]
    assert len(logs) == 1                                     # rank 0 only
    lines = [json.loads(l) for l in open(os.path.join(work, logs[0]))]
    train = [l for l in lines if l["mode"] == "train"]
    assert len(train) == 2 and all(l["loss"] == l["loss"] for l in train)
    assert sum(l["mode"] == "val" for l in lines) == 1       # validation: sample idx on rank idx % 2, one all-reduce of the rows
