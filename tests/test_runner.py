"""The runner and its hooks on the CPU: learning-rate arithmetic against mmcv 0.4.4's
StepLrUpdaterHook / LrUpdaterHook formulas, checkpoint round trip in the reference's file format (optimizer state in
torch.optim.Adam layout), resume, and the TextLoggerHook JSON lines.  The optimizer kernels run under host emulation."""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402
from jperceiver_b200.apis import Runner, StepLrPolicy  # noqa: E402
from jperceiver_b200.apis.runner import load_optimizer_state_dict, optimizer_state_dict  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def emu():
    _lib._handle, _lib._emulated = None, False
    _lib.use_library(build_emulation(), emulated=True)
    yield
    _lib._handle, _lib._emulated = None, False


class Tiny(torch.nn.Module):
    """Stands in for Baseline: forward(inputs) -> (outputs, loss_dict) in training mode."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.conv = torch.nn.Conv2d(3, 4, 3).to(memory_format=torch.channels_last)
        self.fc = torch.nn.Linear(4, 2)

    def forward(self, inputs):
        y = self.conv(inputs["x"]).mean((2, 3))
        z = self.fc(y)
        return {"z": z}, {"a": (z ** 2).mean(), ("b", 0): (y ** 2).mean() * 0.1}


def batches(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"x": torch.randn(2, 3, 8, 8, generator=g)} for _ in range(n)]


def test_step_policy_matches_mmcv_formulas():
    p = StepLrPolicy(1e-4, step=[50])                                   # the reference configs: lr_config(policy='step', step=[50])
    assert p.lr(0, 0) == 1e-4 and p.lr(49, 10 ** 6) == 1e-4
    assert abs(p.lr(50, 0) - 1e-5) < 1e-12 and abs(p.lr(179, 0) - 1e-5) < 1e-12
    q = StepLrPolicy(1e-4, step=[20, 30], gamma=0.5, warmup="linear", warmup_iters=500, warmup_ratio=1.0 / 3)   # the commented-out variant
    assert abs(q.lr(0, 0) - 1e-4 / 3) < 1e-12                            # k = (1 - 0) * (1 - 1/3)
    assert abs(q.lr(0, 250) - 1e-4 * (1 - 0.5 * (2.0 / 3))) < 1e-12
    assert q.lr(0, 500) == 1e-4 and abs(q.lr(25, 10 ** 5) - 5e-5) < 1e-12 and abs(q.lr(30, 10 ** 5) - 2.5e-5) < 1e-12
    r = StepLrPolicy(1.0, step=10, warmup="exp", warmup_iters=4, warmup_ratio=0.1)
    assert abs(r.lr(0, 2) - 0.1 ** 0.5) < 1e-12 and abs(r.lr(25, 100) - 0.01) < 1e-12
    c = StepLrPolicy(1.0, step=10, warmup="constant", warmup_iters=4, warmup_ratio=0.25)
    assert c.lr(0, 3) == 0.25 and c.lr(0, 4) == 1.0
    with pytest.raises(NotImplementedError):
        StepLrPolicy(1.0, step=1, policy="cosine")


def test_runner_trains_logs_checkpoints_and_resumes(tmp_path):
    work = str(tmp_path / "run")
    runner = Runner(Tiny(), dict(type="Adam", lr=1e-2, weight_decay=0), dict(grad_clip=dict(max_norm=35, norm_type=2)), work)
    runner.register_training_hooks(dict(policy="step", warmup=None, step=[1], gamma=0.5), None, dict(interval=1),
                                   dict(interval=2, hooks=[dict(type="TextLoggerHook")]))
    data = batches(4)
    runner.run([data], [("train", 1)], 2)
    assert runner.epoch == 2 and runner.iter == 8
    assert abs(runner.current_lr()[0] - 5e-3) < 1e-12                    # step=[1], gamma=0.5: halved from epoch 1 on
    for name in ("epoch_1.pth", "epoch_2.pth", "latest.pth"):
        assert os.path.exists(os.path.join(work, name))
    lines = [json.loads(l) for l in open(runner.json_log)]
    assert len(lines) == 4 and lines[0]["mode"] == "train" and lines[0]["epoch"] == 1 and lines[0]["iter"] == 2
    assert lines[0]["lr"] == 1e-2 and lines[-1]["lr"] == 5e-3
    assert {"a", "('b', 0)", "loss", "time", "data_time"} <= set(lines[0])
    assert abs(lines[0]["loss"] - (lines[0]["a"] + lines[0]["('b', 0)"])) < 2e-5
    # checkpoint format of the reference: meta / state_dict / optimizer (torch.optim.Adam layout)
    ck = torch.load(os.path.join(work, "epoch_2.pth"), weights_only=False)
    assert set(ck) == {"meta", "state_dict", "optimizer"} and ck["meta"]["epoch"] == 2 and ck["meta"]["iter"] == 8
    assert list(ck["state_dict"]) == list(runner.model.state_dict())
    ref = torch.optim.Adam(Tiny().parameters(), lr=1e-2)
    ref.load_state_dict(ck["optimizer"])                                  # loads into a real torch optimizer
    st = ck["optimizer"]["state"]
    assert len(st) == 4 and float(st[0]["step"]) == 8.0 and st[0]["exp_avg"].shape == runner.model.conv.weight.shape
    # resume: a fresh runner continues bit-for-bit like the original
    cont = Runner(Tiny(), dict(type="Adam", lr=1e-2, weight_decay=0), dict(grad_clip=dict(max_norm=35, norm_type=2)), str(tmp_path / "r2"))
    cont.register_training_hooks(dict(policy="step", warmup=None, step=[1], gamma=0.5), None, None, None)
    cont.resume(os.path.join(work, "latest.pth"))
    assert cont.epoch == 2 and cont.iter == 8
    more = batches(3, seed=7)
    runner.run([more], [("train", 1)], 3)
    cont.run([more], [("train", 1)], 3)
    for (k, a), (_, b) in zip(runner.model.state_dict().items(), cont.model.state_dict().items()):
        assert torch.equal(a, b), k
    a, b = optimizer_state_dict(runner.engine), optimizer_state_dict(cont.engine)
    assert all(torch.equal(a["state"][i]["exp_avg_sq"], b["state"][i]["exp_avg_sq"]) for i in a["state"])


def test_checkpoint_with_module_prefix_loads(tmp_path):
    r = Runner(Tiny(), work_dir=str(tmp_path))
    sd = {"module." + k: v.clone() + 1.0 for k, v in r.model.state_dict().items()}
    path = str(tmp_path / "dp.pth")
    torch.save({"meta": {"epoch": 3, "iter": 30}, "state_dict": sd}, path)
    r.load_checkpoint(path)
    assert torch.allclose(r.model.fc.bias, sd["module.fc.bias"])
    assert r.model.fc.bias.data_ptr() == r.engine.flat._view(r.engine.flat.param, *_view_of(r, "fc.bias")).data_ptr()


def _view_of(r, name):
    names = [n for n, p in r.model.named_parameters() if p.requires_grad]
    i = names.index(name)
    return r.engine.flat.views[i][0], r.engine.flat.params[i]


class TinySnippets(torch.utils.data.Dataset):
    """A map-style dataset with the ``flag`` array the group samplers need (mono_dataset.py sets it to zeros)."""

    def __init__(self, n):
        import numpy as np
        self.flag = np.zeros(n, dtype=np.int64)
        g = torch.Generator().manual_seed(3)
        self.x = torch.randn(n, 3, 8, 8, generator=g)

    def __len__(self):
        return len(self.flag)

    def __getitem__(self, i):
        return {"x": self.x[int(i)], "idx": int(i)}


def test_runner_epochs_follow_the_sampler_plan(tmp_path):
    """``build_dataloader`` + ``Runner.train_epoch``: the epoch is handed to the sampler (DistSamplerSeedHook), each epoch visits
    the plan of ``DistributedGroupSampler`` for that epoch."""
    from jperceiver_b200.datasets import build_dataloader
    ds = TinySnippets(9)
    loader = build_dataloader(ds, 2, 0, dist=True)          # no process group: rank 0 of 1
    seen = []
    model = Tiny()
    fwd = model.forward
    model.forward = lambda inputs: (seen.append(inputs["idx"].tolist()), fwd(inputs))[1]
    runner = Runner(model, dict(type="Adam", lr=1e-3, weight_decay=0), None, str(tmp_path))
    runner.register_training_hooks(dict(policy="step", step=[50]), None, None, None)
    runner.run([loader], [("train", 1)], 2)
    assert runner.iter == 10 and len(seen) == 10            # 9 -> 10 padded snippets, 5 steps per epoch
    for epoch in (0, 1):
        plan = loader.sampler.plan(epoch)[0].reshape(-1, 2).tolist()
        assert seen[5 * epoch:5 * epoch + 5] == plan
    assert seen[:5] != seen[5:]
