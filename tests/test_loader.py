"""Snippet partitioning across ranks: ``jperceiver_b200.datasets.loader`` against index sequences produced by
the reference's own samplers (``oracle/make_golden_sampler.py`` -> ``tests/golden/kat_sampler.json``) — index work, so the bar
is equality — plus size-independent properties at the BASELINE sizes and a 2-process gloo run."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden_sampler import FlagDataset, flags  # noqa: E402  (case definitions only: no reference code runs here)

from jperceiver_b200.datasets import loader as L  # noqa: E402

GOLD = json.load(open(os.path.join(GOLDEN, "kat_sampler.json")))
SAME_RNG = GOLD["torch"].split("+")[0] == torch.__version__.split("+")[0] and GOLD["numpy"] == np.__version__
needs_same_rng = pytest.mark.skipif(not SAME_RNG, reason="golden sequences were drawn with torch %s / numpy %s" % (GOLD["torch"], GOLD["numpy"]))


@needs_same_rng
@pytest.mark.parametrize("case", GOLD["dgs"], ids=lambda c: "%s-n%d-spg%d-w%d-e%d" % (c["flag"], c["n"], c["spg"], c["world"], c["epoch"]))
def test_distributed_group_sampler_equals_reference(case):
    ds = FlagDataset(flags(case["flag"], case["n"]))
    for r in range(case["world"]):
        s = L.DistributedGroupSampler(ds, case["spg"], case["world"], r)
        s.set_epoch(case["epoch"])
        assert len(s) == case["len"]
        assert list(s) == case["indices"][r]
    assert s.plan().tolist() == case["indices"]          # the whole job's plan, from any rank


@needs_same_rng
def test_distributed_and_group_sampler_equal_reference():
    for case in GOLD["ds"]:
        ds = FlagDataset(np.zeros(case["n"]))
        for r in range(case["world"]):
            s = L.DistributedSampler(ds, case["world"], r, shuffle=case["shuffle"])
            s.set_epoch(case["epoch"])
            assert len(s) == case["len"] and list(s) == case["indices"][r]
    for case in GOLD["gs"]:
        ds = FlagDataset(flags(case["flag"], case["n"]))
        np.random.seed(case["seed"])
        s = L.GroupSampler(ds, case["spg"])
        assert len(s) == case["len"] and [int(i) for i in s] == case["indices"]


@pytest.mark.parametrize("n,spg,world", [(40109, 4, 1), (40109, 8, 8), (23488, 3, 8), (7, 2, 2)])
def test_partition_properties(n, spg, world):
    """KITTI-odometry / raw sized splits at the BASELINE batch shapes (4 per GPU x 1, 8 per GPU x 8, global batch 24 over 8)."""
    flag = flags("two", n)
    s = L.DistributedGroupSampler(FlagDataset(flag), spg, world, 0)
    for epoch in (0, 1):
        plan = s.plan(epoch)
        assert plan.shape == (world, len(s)) and len(s) % spg == 0
        everything = plan.reshape(-1)
        assert set(everything.tolist()) == set(range(n))                  # every snippet is seen
        counts = np.bincount(everything, minlength=n)
        assert counts.max() <= 2 and counts.sum() - n < 2 * spg * world      # padding: at most one wrap per group
        chunks = flag[plan.reshape(-1, spg)]
        assert (chunks == chunks[:, :1]).all()                            # a step's samples share one group
    assert not np.array_equal(s.plan(0), s.plan(1)) or n < 4
    assert np.array_equal(s.plan(3), L.DistributedGroupSampler(FlagDataset(flag), spg, world, world - 1).plan(3))


def test_group_too_small_to_pad_raises_like_the_reference():
    with pytest.raises(AssertionError):          # sampler.py:142 — one wrap cannot fill 3 -> 8
        list(L.DistributedGroupSampler(FlagDataset(np.zeros(3)), 2, 4, 0))


class ToySnippets(torch.utils.data.Dataset):
    """Samples shaped like MonoDataset's: a dict keyed by tuples."""

    def __init__(self, n):
        self.flag = np.zeros(n, dtype=np.int64)

    def __len__(self):
        return len(self.flag)

    def __getitem__(self, i):
        i = int(i)                    # GroupSampler yields 0-d LongTensors, as the reference's does
        return {("color", 0, 0): torch.full((3, 4, 6), float(i)), ("K", 0): np.eye(4, dtype=np.float32) * i, "idx": i}


def test_build_dataloader_single_process():
    ds = ToySnippets(10)
    np.random.seed(0)
    dl = L.build_dataloader(ds, 3, 0, num_gpus=1, dist=False)
    batches = list(dl)
    assert len(batches) == 4                                   # 12 padded indices / 3 (drop_last=True)
    b = batches[0]
    assert b[("color", 0, 0)].shape == (3, 3, 4, 6) and b[("K", 0)].shape == (3, 4, 4) and b["idx"].shape == (3,)
    assert torch.equal(b[("color", 0, 0)][:, 0, 0, 0].long(), b["idx"])
    dl = L.build_dataloader(ds, 2, 0, dist=True, shuffle=False)   # not initialised: rank 0 of 1, sequential order
    assert [b["idx"].tolist() for b in dl] == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]]
    assert isinstance(L.build_dataloader(ds, 2, 0, dist=True).sampler, L.DistributedGroupSampler)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ds = ToySnippets(21)
        dl = L.build_dataloader(ds, 2, 0, dist=True)                 # rank / world from the process group
        seen = []
        for epoch in (0, 1):
            dl.sampler.set_epoch(epoch)
            mine = torch.cat([b["idx"] for b in dl])
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            seen.append(torch.stack(gathered))
        if rank == 0:
            torch.save(seen, out)
    finally:
        dist.destroy_process_group()


def test_two_ranks_take_disjoint_slices_of_one_plan_gloo(tmp_path):
    out = str(tmp_path / "seen.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    seen = torch.load(out)
    ref = L.DistributedGroupSampler(ToySnippets(21), 2, 2, 0)
    for epoch, got in enumerate(seen):
        assert got.shape == (2, 12)                                  # 21 -> 24 padded, 12 per rank
        assert np.array_equal(got.numpy(), ref.plan(epoch))
        assert set(got.reshape(-1).tolist()) == set(range(21))


def _grads_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mono.core import DistOptimizerHook, allreduce_grads        # the reference's import path (mono/core/__init__.py:6)
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
        net[0].bias.requires_grad_(False)                                # a parameter without gradient is skipped, as in the reference
        x = torch.randn(6, 5, generator=torch.Generator().manual_seed(10 + rank))
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in net.parameters() if p.grad is not None]
        gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
        for g, buf in zip(local, gathered):
            dist.all_gather(buf, g)
        expect = [sum(buf) / world for buf in gathered]
        for coalesce in (True, False):
            for p, g in zip([p for p in net.parameters() if p.grad is not None], local):
                p.grad.copy_(g)
            allreduce_grads(net, coalesce=coalesce, bucket_size_mb=25)
            got = [p.grad for p in net.parameters() if p.grad is not None]
            assert all((a - b).abs().max().item() < 1e-6 for a, b in zip(got, expect)), coalesce
        hook = DistOptimizerHook(grad_clip=dict(max_norm=35, norm_type=2))
        assert hook.grad_clip["max_norm"] == 35 and hook.coalesce and hook.bucket_size_mb == -1
        if rank == 0:
            torch.save(True, out)
    finally:
        dist.destroy_process_group()


def test_allreduce_grads_is_the_mean_over_ranks_gloo(tmp_path):
    """``mono.core.utils.allreduce_grads`` (dist_utils.py:34-44): every ``param.grad`` becomes the mean over ranks, coalesced or not."""
    out = str(tmp_path / "ok.pt")
    mp.spawn(_grads_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out) is True
