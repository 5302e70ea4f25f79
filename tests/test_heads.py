"""Parity of the small fused operators around the trunks (csrc/heads.cu) with the plain PyTorch formulation of the
reference lines they replace: input prologue (ResnetEncoder.py:99, net.py:633-638), dropout (depth_decoder.py:47-48),
pose head (pose_decoder.py:22-26, net.py:704-756) incl. its hand-derived backward; and the KAT2 vector of SURVEY.md §8c.
``[emu]`` = host emulation of the same sources (CPU container), ``[gpu]`` = the C ABI on cuda:0."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402
from emu.torch_ops import cct_attention_torch, pose_head_torch  # noqa: E402

from jperceiver_b200 import _lib, netops as ops  # noqa: E402


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


@pytest.mark.parametrize("pair,out_hw", [(False, None), (True, (24, 40)), (False, (32, 32)), (True, None)])
def test_image_prep_vs_torch(dev, pair, out_hw):
    g = torch.Generator().manual_seed(0)
    ims = [torch.rand(2, 3, 20, 36, generator=g) for _ in range(2 if pair else 1)]
    got = ops.image_prep([im.to(dev) for im in ims], out_hw)
    ref = []
    for im in ims:
        if out_hw is not None:
            im = F.interpolate(im, list(out_hw), mode="bilinear", align_corners=False)
        ref.append((im - 0.45) / 0.225)
    ref = torch.cat(ref, 1)
    C = ref.shape[1]
    assert got.shape[1] == (8 if pair else 4) and got.is_contiguous(memory_format=torch.channels_last)
    assert (got[:, :C].cpu() - ref).abs().max().item() < 2e-6       # bilinear weights in a different association order
    assert got[:, C:].abs().max().item() == 0.0                      # zero channel padding


def test_dropout_mask_and_counter_draw(dev):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 8, 6, 10, generator=g).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    mask = (torch.rand(2, 8, 6, 10, generator=g) >= 0.5).float().to(dev)
    y = ops.dropout(x, 0.5, True, mask)
    assert (y - x * mask * 2.0).abs().max().item() == 0.0
    gy = torch.randn(y.shape, generator=g).to(dev)
    (gx,) = torch.autograd.grad(y, x, gy)
    assert (gx - gy * mask * 2.0).abs().max().item() == 0.0
    # counter-based draw: every element is 0 or 2x, about half are kept, the backward regenerates the same mask
    big = torch.ones(4, 64, 16, 16, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yb = ops.dropout(big, 0.5, True)
    vals = torch.unique(yb.detach().cpu())
    assert set(vals.tolist()) <= {0.0, 2.0}
    keep = (yb.detach() > 0).float().mean().item()
    assert 0.47 < keep < 0.53
    (gb,) = torch.autograd.grad(yb, big, torch.ones_like(yb))
    assert (gb - yb.detach()).abs().max().item() == 0.0
    assert ops.dropout(big, 0.5, False) is big and ops.dropout(big, 0.0, True) is big


@pytest.mark.parametrize("invert", [False, True])
def test_pose_head_forward_backward_vs_torch(dev, invert):
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(3, 6, 6, 20, generator=g) * 3 + 1.0)
    x0 = x.clone().requires_grad_(True)
    ref = pose_head_torch(x0, invert)
    G = torch.randn(3, 4, 4, generator=g)
    (g0,) = torch.autograd.grad(ref, x0, G)
    x1 = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = ops.pose_head(x1, invert)
    assert (got.cpu() - ref).abs().max().item() < 2e-6
    (g1,) = torch.autograd.grad(got, x1, G.to(dev))
    assert (g1.cpu() - g0).abs().max().item() <= 1e-5 * g0.abs().max().item() + 1e-12


def test_pose_head_kat2(dev):
    """transformation_from_parameters(aa=[.01,-.02,.03], t=[.1,-.05,.2]) — feed a constant map whose
    mean x 0.01 equals those parameters."""
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    v = torch.tensor([.01, -.02, .03, .1, -.05, .2]) * 100.0
    x = v.view(1, 6, 1, 1).expand(1, 6, 4, 5).contiguous().to(dev)
    fwd = ops.pose_head(x, False).cpu().numpy()[0]
    inv = ops.pose_head(x, True).cpu().numpy()[0]
    assert np.abs(fwd - kat["kat2_fwd"].reshape(4, 4)).max() < 2e-6
    assert np.abs(inv - kat["kat2_inv"].reshape(4, 4)).max() < 2e-6


def test_cvp_mlp_forward_backward_vs_torch(dev):
    g = torch.Generator().manual_seed(3)
    B, C, h = 2, 12, 4
    n = h * h
    x = torch.randn(B, C, h, h, generator=g)
    fc0, fc2 = torch.nn.Linear(n, n), torch.nn.Linear(n, n)
    with torch.no_grad():
        for p in list(fc0.parameters()) + list(fc2.parameters()):
            p.copy_(torch.randn(p.shape, generator=g) * 0.3)
    x0 = x.clone().requires_grad_(True)
    ref = F.relu(F.linear(F.relu(F.linear(x0.reshape(B, C, n), fc0.weight, fc0.bias)), fc2.weight, fc2.bias)).reshape(B, C, h, h)
    G = torch.randn(ref.shape, generator=g)
    gref = torch.autograd.grad(ref, [x0, fc0.weight, fc0.bias, fc2.weight, fc2.bias], G)
    import copy
    f0, f2 = copy.deepcopy(fc0).to(dev), copy.deepcopy(fc2).to(dev)
    x1 = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = ops.cvp_mlp(x1, f0, f2)
    assert (got.cpu() - ref).abs().max().item() < 1e-5
    ggot = torch.autograd.grad(got, [x1, f0.weight, f0.bias, f2.weight, f2.bias], G.to(dev))
    for a, b in zip(gref, ggot):
        assert (a - b.cpu()).abs().max().item() <= 1e-5 * max(a.abs().max().item(), 1.0)


class _P:   # parameter holder with the attribute names CrossViewTransformer uses
    pass


def test_cct_attention_forward_backward_vs_torch(dev):
    """The fused cross-view-transformer core against the library formulation (bmm / max / gather / broadcast matmul) with the
    same tensor-core convolutions around it (host emulation runs those through the library convolution)."""
    g = torch.Generator().manual_seed(4)
    B, C, h = 2, 16, 4
    mk = lambda *s: torch.randn(*s, generator=g)
    tensors = [mk(B, C, h, h) for _ in range(4)]
    p = _P()
    for name, co, ci, k in (("query_conv", C // 8, C, 1), ("key_conv", C // 8, C, 1), ("value_conv", C, C, 1), ("f_conv", C, 2 * C, 3),
                            ("query_conv_depth", C // 8, C, 1), ("key_conv_depth", C // 8, C, 1), ("value_conv_depth", C, C, 1)):
        m = torch.nn.Conv2d(ci, co, k)
        with torch.no_grad():
            m.weight.copy_(mk(*m.weight.shape) * 0.3); m.bias.copy_(mk(*m.bias.shape) * 0.1)
        setattr(p, name, m.to(dev).to(memory_format=torch.channels_last))
    res = {}
    for backend in ("torch", "jpb"):
        ins = [t.clone().to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) for t in tensors]
        out, S, attn = ops.cct_attention(*ins, p) if backend == "jpb" else cct_attention_torch(*ins, p, ops.conv2d)
        G = torch.Generator().manual_seed(9)
        go = torch.randn(out.shape, generator=G).to(dev)
        loss = (out * go).sum() + (S * 0.7).sum() - (attn * 0.3).sum()
        params = [getattr(p, n_).weight for n_ in ("query_conv", "key_conv", "value_conv", "f_conv", "key_conv_depth", "value_conv_depth")]
        grads = torch.autograd.grad(loss, ins + params)
        res[backend] = [out.detach().cpu(), S.detach().cpu(), attn.detach().cpu()] + [t.cpu() for t in grads]
    for a, b in zip(res["torch"], res["jpb"]):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() <= 2e-4 * max(a.abs().max().item(), 1.0)
