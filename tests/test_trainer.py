"""Host-side training logic: flat parameter/gradient buffers, the fused clip+Adam step against
torch.optim.Adam + clip_grad_norm_, the reference's loss-summing contract, config loading, and the
data-parallel gradient exchange on a 2-process gloo group (CPU, host emulation of the kernels)."""
import os
import sys

import pytest
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402
from jperceiver_b200.apis import Config, TrainEngine, build_optimizer  # noqa: E402
from jperceiver_b200.apis.trainer import FlatParameters, loss_scalars  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def emulation():
    _lib.use_library(build_emulation(), emulated=True)
    yield
    _lib._handle, _lib._emulated = None, False


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 8, 3, padding=1, bias=False)   # a bias in front of BN has a pure-noise gradient, which Adam amplifies
        self.conv.weight.data = self.conv.weight.data.contiguous(memory_format=torch.channels_last)
        self.bn = nn.BatchNorm2d(8)
        self.fc = nn.Linear(8, 4)
        self.unused = nn.Linear(4, 4)   # never receives a gradient (like the ResNet fc / res_conv of the real model)

    def forward(self, data):
        x = data["x"]
        y = self.fc(torch.relu(self.bn(self.conv(x))).mean((2, 3)))
        return {"y": y}, {"a": (y ** 2).mean(), ("b", 0): y.abs().mean() * 0.5}


def test_loss_sum_counts_every_entry():
    names, vals, total = loss_scalars({"topview_loss": torch.tensor(2.0), "layout_loss": torch.tensor(3.0), ("smooth_loss", 0): torch.tensor(0.5)})
    assert names == ["topview_loss", "layout_loss", "('smooth_loss', 0)"] and total.item() == 5.5
    with pytest.raises(TypeError):
        loss_scalars({"bad": 1.0})


def test_flat_views_keep_layout_and_values():
    m = Tiny()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = FlatParameters(m)
    assert flat.numel == sum((p.numel() + 63) // 64 * 64 for p in m.parameters())
    assert all((p.data_ptr() - flat.param.data_ptr()) % 256 == 0 for p in m.parameters())
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    assert m.conv.weight.is_contiguous(memory_format=torch.channels_last)
    assert m.conv.weight.data_ptr() == flat.param.data_ptr()
    flat.param.zero_()
    assert m.fc.weight.abs().sum().item() == 0.0


@pytest.mark.parametrize("max_norm", [None, 0.05])
def test_fused_adam_matches_torch(max_norm):
    torch.manual_seed(0)
    a, b = Tiny(), Tiny()
    b.load_state_dict(a.state_dict())
    opt_ref = torch.optim.Adam(a.parameters(), lr=1e-2, weight_decay=0)
    eng = TrainEngine(b, dict(type="Adam", lr=1e-2, weight_decay=0), dict(max_norm=max_norm, norm_type=2) if max_norm else {})
    for it in range(4):
        x = torch.randn(4, 3, 8, 8, generator=torch.Generator().manual_seed(it))
        opt_ref.zero_grad()
        _, la = a({"x": x})
        sum(la.values()).backward()
        if max_norm:
            torch.nn.utils.clip_grad_norm_(a.parameters(), max_norm)
        opt_ref.step()
        eng.step({"x": x})
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa - pb).abs().max().item() <= 2e-6 + 1e-5 * pa.abs().max().item(), k
    assert int(eng.optimizer.step_count.item()) == 4


def test_reference_configs_load_unchanged():
    cfg_dir = "/root/reference/config"
    if not os.path.isdir(cfg_dir):
        pytest.skip("reference tree not present")
    for name in ("cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20_B1", "cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20",
                 "cfg_kitti_baseline_raw_boundary_ce_iou_1024_20", "cfg_kitti_baseline_argo_both_boundary_ce_iou_1024_20_B1",
                 "cfg_kitti_baseline_kitti_odom_8pugsB24_lr1e-4_ce_eigen"):
        cfg = Config.fromfile(os.path.join(cfg_dir, name + ".py"))
        assert cfg.model.name == "Baseline" and cfg.model["type"] == cfg.model.type
        assert cfg.optimizer.type == "Adam" and cfg.optimizer_config.grad_clip.max_norm == 35


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import build_emulation as be
    _lib.use_library(be(), emulated=True)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = Tiny()
    eng = TrainEngine(m, dict(type="Adam", lr=1e-2, weight_decay=0), dict(max_norm=35, norm_type=2))
    x = torch.randn(4, 3, 8, 8, generator=torch.Generator().manual_seed(100 + rank))
    # expected: average over ranks of the single-process gradients (BN statistics stay per-rank)
    eng.flat.zero_grad()
    _, l = m({"x": x})
    sum(l.values()).backward()
    local = eng.flat.grad.clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    eng.exchange_gradients()
    got = eng.flat.grad / world
    ok_grad = bool((got - expect).abs().max().item() < 1e-6)
    eng.step({"x": x})
    params = eng.flat.param.clone()
    allp = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(allp, params)
    ok_sync = bool(all(torch.equal(allp[0], p) for p in allp))
    q.put((rank, ok_grad, ok_sync))
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


@pytest.mark.gpu
def test_direct_gradient_accumulation_matches_plain_autograd_gpu():
    """The training-step backward adds parameter gradients straight into ``param.grad`` (views of the flat gradient buffer) and
    hands autograd None.  Per operator, on identical inputs, that must equal the plain path (gradient tensors returned to
    autograd) — and a second backward must accumulate (2x), as parameters shared between calls (the pose encoder runs once per
    source frame) require.  Whole-network comparisons cannot check this: under TF32 operand rounding the network amplifies
    fp32 atomic-order noise to 1e-4..1e-3 per activation, so two identical runs already differ by a few percent in gradient
    norm (tools/diag_noise.py)."""
    import torch
    from jperceiver_b200 import _lib, functional as JF, netops as ops
    _lib._handle, _lib._emulated = None, False
    _lib.lib()
    dev = torch.device("cuda:0")
    CL = torch.channels_last
    g = torch.Generator().manual_seed(21)

    def check(build, params):
        """build() -> scalar loss using ``params`` (leaf tensors requiring grad)."""
        plain = torch.autograd.grad(build(), params)
        for p in params:
            p.grad = torch.zeros_like(p)          # same strides as the parameter (channels-last weights stay channels-last)
        for rep in (1, 2):
            JF.DIRECT_GRAD = True
            try:
                build().backward()
            finally:
                JF.DIRECT_GRAD = False
            for p, ref in zip(params, plain):
                err = (p.grad - rep * ref).abs().max().item()
                assert err <= 2e-5 * rep * max(ref.abs().max().item(), 1e-6), (rep, tuple(p.shape), err)

    # convolution: weight gradient + bias gradient, 3x3 with activation and 1x1 without
    for cin, cout, k, act in ((64, 128, 3, "leaky"), (128, 256, 1, "none"), (256, 16, 3, "relu")):
        x = torch.randn(2, cin, 24, 40, generator=g).to(dev).contiguous(memory_format=CL)
        w = (torch.randn(cout, cin, k, k, generator=g) * 0.05).to(dev).contiguous(memory_format=CL).requires_grad_(True)
        b = torch.randn(cout, generator=g).to(dev).requires_grad_(True)
        G = torch.randn(2, cout, 24, 40, generator=g).to(dev)
        check(lambda: (ops.conv2d(x, w, b, pad=k // 2, act=act) * G).sum(), [w, b])
    # BatchNorm affine parameters
    bn = torch.nn.BatchNorm2d(64).to(dev).train()
    xb = torch.randn(2, 64, 24, 40, generator=g).to(dev).contiguous(memory_format=CL)
    Gb = torch.randn(2, 64, 24, 40, generator=g).to(dev)
    check(lambda: (ops.batchnorm(xb, bn, True, relu=True) * Gb).sum(), [bn.weight, bn.bias])
    # CVP transform module
    fc0, fc2 = torch.nn.Linear(16, 16).to(dev), torch.nn.Linear(16, 16).to(dev)
    xc = torch.randn(2, 32, 4, 4, generator=g).to(dev).contiguous(memory_format=CL)
    Gc = torch.randn(2, 32, 4, 4, generator=g).to(dev)
    check(lambda: (ops.cvp_mlp(xc, fc0, fc2) * Gc).sum(), [fc0.weight, fc0.bias, fc2.weight, fc2.bias])


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import build_emulation as be
    _lib.use_library(be(), emulated=True)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    engines = []
    for overlap in ("1", "0"):
        os.environ["JPB_OVERLAP_ALLREDUCE"] = overlap
        torch.manual_seed(0)
        m = Tiny()
        eng = TrainEngine(m, dict(type="Adam", lr=1e-2, weight_decay=0), dict(max_norm=35, norm_type=2))
        if eng.exchange is not None:
            eng.exchange.bucket_bytes, eng.exchange.tail_bytes = 512, 128      # several buckets on a model this small
        engines.append(eng)
    for it in range(4):
        x = torch.randn(4, 3, 8, 8, generator=torch.Generator().manual_seed(100 * it + rank))
        for eng in engines:
            eng.step({"x": x})
    a, b = engines
    ok_mode = a.exchange is not None and a.exchange.mode == "run" and len(a.exchange.buckets) >= 3 and b.exchange is None
    # the same parameter values whichever way the gradient travelled (the layouts differ: compare per parameter)
    ok_same = all(torch.equal(pa, pb) for pa, pb in zip(a.model.parameters(), b.model.parameters()))
    flat = torch.cat([p.detach().reshape(-1) for p in a.model.parameters()])
    allp = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    ok_sync = all(torch.equal(allp[0], p) for p in allp)
    covered = sorted((lo, hi) for lo, hi, _ in a.exchange.buckets)
    ok_cover = covered[0][0] == 0 and covered[-1][1] == a.flat.numel and all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    q.put((rank, bool(ok_mode), bool(ok_same), bool(ok_sync), bool(ok_cover)))
    dist.destroy_process_group()


def test_overlapped_bucketed_allreduce_matches_single_exchange_gloo_world2():
    """GradExchange: step 1 traces the gradient completion order and re-lays the flat buffers; steps 2.. all-reduce bucket by
    bucket during backward.  Four steps on two gloo ranks must give bit-identical parameters to the one-piece exchange, identical
    across ranks, and the buckets must tile the flat buffer exactly."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 2000
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(all(r[1:]) for r in res), res
