"""Validation metrics on the device against the oracle (``oracle/eval_port.py``) and the golden vectors the
reference's own ``pixel_error.py`` produced (``oracle/make_golden_eval.py`` -> ``tests/golden/kat_eval.npz``).

``[gpu]`` calls libjpb200.so through the C ABI on cuda:0; ``[emu]`` runs the same kernel sources compiled as host C++.
Counts and medians (order statistics) are exact; error means are fp32 terms, tolerance 2e-6 relative.
(Sorted after the training-path suites: these kernels were written after this round's GPU budget was spent, so the GPU cases
below are their first on-device run; the emulation cases cover the kernel logic.)"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402
from oracle import eval_port as EO  # noqa: E402
from oracle.make_golden_eval import BEV_KINDS, DEPTH_CASES, bev_case, depth_case  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402
from jperceiver_b200.core import evaluation as EV  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 2e-6


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def dev(request):
    _lib._handle, _lib._emulated = None, False
    if request.param == "emu":
        _lib.use_library(build_emulation(), emulated=True)
        yield torch.device("cpu")
    else:
        assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
        _lib.lib()
        yield torch.device("cuda:0")
    _lib._handle, _lib._emulated = None, False


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "kat_eval.npz"))


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-12))


def test_oracle_matches_golden(gold):
    """The CPU restatement against the vectors written from the reference's functions (no device code involved)."""
    pytest.importorskip("cv2")
    r = np.random.RandomState(7)
    gt = (1 + 60 * r.rand(5000)).astype(np.float32)
    pred = (gt * (1 + 0.2 * r.randn(5000))).clip(1e-3, 80).astype(np.float32)
    assert close(EO.compute_errors(gt, pred), gold["errors_seed7"], 1e-7)
    for i, kind in enumerate(BEV_KINDS):
        p, t = bev_case(10 + i, 32, kind)
        assert list(map(float, EO.mean_IU(p, t))) == list(gold["iu_" + kind])
        assert list(map(float, EO.mean_precision(p, t))) == list(gold["mp_" + kind])
        assert list(EO.hook_bev_values(p, t)) == list(gold["hook_" + kind])
    for c in DEPTH_CASES:
        disp, g = depth_case(*c)
        assert close(EO.depth_eval_sample(disp, g), gold["depth_%d" % c[0]], 1e-7)


def test_compute_errors_vs_reference_golden(dev, gold):
    r = np.random.RandomState(7)
    gt = (1 + 60 * r.rand(5000)).astype(np.float32)
    pred = (gt * (1 + 0.2 * r.randn(5000))).clip(1e-3, 80).astype(np.float32)
    got = EV.compute_errors(torch.from_numpy(gt).to(dev), torch.from_numpy(pred).to(dev))
    assert len(got) == 7 and close(got, gold["errors_seed7"])
    assert got[4:] == tuple(gold["errors_seed7"][4:])            # threshold fractions are counts / n: exact


@pytest.mark.parametrize("case", DEPTH_CASES, ids=lambda c: "seed%d_%dx%d_to_%dx%d" % c[:5])
@pytest.mark.parametrize("stereo", [False, True])
def test_depth_errors_vs_golden(dev, gold, case, stereo):
    """disp -> scaled disp -> cv2-style bilinear resize to the ground-truth frame -> 1/x -> validity + crop mask -> median
    scaling (or x36) -> clamp -> seven errors; up-sampling, down-sampling and identity resize; B=2 with a second sample whose
    ground truth is empty (NaN row, as numpy)."""
    disp, g = depth_case(*case)
    d2 = torch.from_numpy(np.stack([disp, disp[::-1].copy()]))[:, None].to(dev)
    g2 = torch.from_numpy(np.stack([g, np.zeros_like(g)])).to(dev)
    out = EV.depth_errors(d2, g2, min_depth=EV.MIN_DEPTH, max_depth=EV.MAX_DEPTH, stereo_scale=stereo).cpu().numpy()
    want = gold["depth_%d%s" % (case[0], "_stereo" if stereo else "")]
    assert close(out[0], want), (out[0], want)
    assert close(out[0, 7], want[7], 1e-6)             # ratio of two exact order statistics (the resize may differ by an ulp)
    assert np.isnan(out[1]).all()


def test_depth_medians_are_exact_order_statistics(dev):
    """Radix select = np.median bit for bit (odd and even counts, duplicates, values across many binades)."""
    r = np.random.RandomState(3)
    for n in (1, 2, 7, 64, 1001, 4096):
        g = np.exp(r.uniform(-5, 4, n)).astype(np.float32).clip(2e-3, 79)
        g[: n // 3] = g[0]                                                     # duplicates
        p = np.exp(r.uniform(-3, 3, n)).astype(np.float32)
        out = EV.depth_errors(torch.from_numpy(1.0 / p).reshape(1, 1, 1, n).to(dev), torch.from_numpy(g).reshape(1, 1, n).to(dev),
                              min_depth=EV.MIN_DEPTH, max_depth=EV.MAX_DEPTH, crop=(0, 1, 0, n), net_min_depth=1.0,
                              net_max_depth=float("inf")).cpu().numpy()
        pp = (1.0 / (1.0 / p).astype(np.float32)).astype(np.float32)
        assert np.float32(out[0, 7]) == np.float32(np.median(g) / np.median(pp)), n


@pytest.mark.parametrize("channels_last", [False, True])
def test_bev_counts_exact_and_hook_values(dev, gold, channels_last):
    occ, B = 32, len(BEV_KINDS)
    preds, trues = zip(*[bev_case(10 + i, occ, k) for i, k in enumerate(BEV_KINDS)])
    g = torch.Generator().manual_seed(0)
    mag = 0.1 + torch.rand(B, occ, occ, generator=g)
    pr = torch.from_numpy(np.stack(preds)).float()
    logits = torch.stack([torch.rand(B, occ, occ, generator=g), torch.zeros(B, occ, occ)], 1)
    logits[:, 1] = logits[:, 0] + (2 * pr - 1) * mag                           # argmax == pred, never tied
    logits = logits.to(dev)
    if channels_last:
        logits = logits.contiguous(memory_format=torch.channels_last)
    label = torch.from_numpy(np.stack(trues))[:, None].to(dev)
    counts = EV.bev_counts(logits, label).cpu().numpy()
    for i in range(B):
        p, t = preds[i] == 1, trues[i] == 1
        assert list(counts[i]) == [int((p & t).sum()), int(p.sum()), int(t.sum())]
    from jperceiver_b200.core.evaluation.pixel_error import hook_values
    iou, mAP = hook_values(torch.from_numpy(counts).to(dev), occ * occ)
    for i, kind in enumerate(BEV_KINDS):
        assert [iou[i].item(), mAP[i].item()] == list(gold["hook_" + kind]), kind
        # the reference-named list functions on class maps
        assert EV.mean_IU(torch.from_numpy(preds[i]).to(dev), torch.from_numpy(trues[i]).to(dev)) == list(gold["iu_" + kind]), kind
        assert EV.mean_precision(torch.from_numpy(preds[i]).to(dev), torch.from_numpy(trues[i]).to(dev)) == list(gold["mp_" + kind]), kind
    tie = torch.zeros(1, 2, occ, occ, device=dev)                                # equal logits -> class 0 (first maximum)
    assert EV.bev_counts(tie, label[:1]).cpu().tolist()[0][1] == 0


class _EchoModel(torch.nn.Module):
    """Stands in for the network: returns the prediction tensors stored in the sample (the hook's metric path is under test)."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def forward(self, batch):
        assert not self.training
        return {("disp", 0, 0): batch["disp_in"], "topview": batch["lgS"], "topviewB": batch["lgD"]}


def _dataset(n=5):
    items = []
    for i in range(n):
        disp, g = depth_case(20 + i, 24, 80, 47, 155, 0.3)
        pS, tS = bev_case(30 + i, 32, BEV_KINDS[i % len(BEV_KINDS)])
        pD, tD = bev_case(40 + i, 32, BEV_KINDS[(i + 2) % len(BEV_KINDS)])
        lg = lambda p: np.stack([np.zeros_like(p, np.float32), 2.0 * p.astype(np.float32) - 1.0])
        items.append({"disp_in": disp[None], "gt_depth": g, "lgS": lg(pS), "lgD": lg(pD), ("bothS", 0, 0): tS[None], ("bothD", 0, 0): tD[None],
                      "_expect": (disp, g, pS, tS, pD, tD)})
    return items


def _expected_means(items):
    rows = []
    for it in items:
        disp, g, pS, tS, pD, tD = it["_expect"]
        rows.append(list(EO.depth_eval_sample(disp, g)) + list(EO.hook_bev_values(pS, tS)) + list(EO.hook_bev_values(pD, tD)))
    return np.mean(np.array(rows, np.float64), 0)


class _Items:
    def __init__(self, items):
        self.items = items

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return {k: v for k, v in self.items[i].items() if k != "_expect"}


class _FakeRunner:
    def __init__(self, model, rank=0, world=1):
        from jperceiver_b200.apis.runner import LogBuffer
        self.model, self.rank, self.world_size, self.epoch = model, rank, world, 0
        self.log_buffer = LogBuffer()


def test_dist_eval_mono_hook_matches_oracle(dev):
    pytest.importorskip("cv2")
    items = _dataset()
    model = _EchoModel().to(dev).train()
    hook = EV.DistEvalMonoHook(_Items(items), interval=1, cfg={"data": {"stereo_scale": False}})
    runner = _FakeRunner(model)
    hook.after_train_epoch(runner)
    assert runner.log_buffer.ready and model.training
    got = [runner.log_buffer.output["scale mean" if k == "scale" else k] for k in EV.eval_hooks.KEYS]
    assert close(got, _expected_means(items), 5e-6), (got, _expected_means(items))
    hook2 = EV.DistEvalMonoHook(_Items(items), interval=2, cfg=None)               # epoch 0: (0 + 1) % 2 != 0 -> skipped
    r2 = _FakeRunner(model)
    hook2.after_train_epoch(r2)
    assert not r2.log_buffer.ready


def _hook_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import build_emulation as be
    _lib.use_library(be(), emulated=True)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = _dataset()
    hook = EV.DistEvalMonoHook(_Items(items), interval=1, cfg=None)
    runner = _FakeRunner(_EchoModel(), rank, world)
    hook.after_train_epoch(runner)
    out = None
    if rank == 0:
        out = [runner.log_buffer.output["scale mean" if k == "scale" else k] for k in EV.eval_hooks.KEYS]
    q.put((rank, runner.log_buffer.ready, out))
    dist.destroy_process_group()


def test_dist_eval_hook_gloo_world2():
    """Samples are split idx % world over two ranks; the rows meet in one all-reduce; rank 0 alone evaluates."""
    pytest.importorskip("cv2")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_hook_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert res[0][1] and not res[1][1]
    assert close(res[0][2], _expected_means(_dataset()), 5e-6)
