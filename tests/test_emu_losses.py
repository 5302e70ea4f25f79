"""Kernel-logic checks in the GPU-less container: the CUDA-core kernels compiled as host C++
(tests/emu) against the oracle port and its autograd gradients.  The same comparisons run on the
real library in tests/test_gpu_losses.py."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pat

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402
from oracle import port as O  # noqa: E402

from jperceiver_b200 import _lib, functional as JF  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def emulation():
    _lib.use_library(build_emulation(), emulated=True)
    yield
    _lib._handle, _lib._emulated = None, False


def _photo_case(B=2, H=24, W=40, s=0, F=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    hs, ws = H >> (s + 1), W >> (s + 1)
    base = torch.rand(B, 3, H // 4 + 2, W // 4 + 2, generator=g)
    up = torch.nn.functional.interpolate(base, (H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    target = (0.8 * up + 0.2 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
    sources = [(0.8 * torch.roll(up, (f, 2 * f), (2, 3)) + 0.2 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
               for f in range(1, F + 1)]
    disp = (0.05 + 0.9 * torch.rand(B, 1, hs, ws, generator=g))
    K = torch.tensor([[.58 * W, 0, .5 * W, 0], [0, 1.92 * H, .5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).repeat(B, 1, 1)
    invK = torch.linalg.pinv(K)
    Ts = []
    for f in range(F):
        aa = 0.02 * torch.randn(B, 3, generator=g)
        t = 0.1 * torch.randn(B, 3, generator=g)
        Ts.append(O.pose_matrix(aa, t, invert=(f == 0)))
    return target, sources, disp, K, invK, Ts


def test_photometric_kat5():
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    H, W = 8, 12
    K = torch.tensor([[.58 * W, 0, .5 * W, 0], [0, 1.92 * H, .5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).unsqueeze(0)
    Tf, Ti = torch.from_numpy(kat["kat2_fwd"]), torch.from_numpy(kat["kat2_inv"])
    loss, winner, idx, warped = JF.photometric_loss(
        0.1 + 0.8 * pat((1, 1, 4, 6), 8), pat((1, 3, 8, 12), 2), [pat((1, 3, 8, 12), 6), pat((1, 3, 8, 12), 7)],
        [Ti, Tf], K, torch.linalg.pinv(K), num_scales=1, noise_scale=0.0, debug_outputs=True)
    assert abs(loss.item() - float(kat["kat5_mean"])) < 2e-6
    assert list(np.bincount(idx.flatten().numpy(), minlength=4)) == [0, 0, 41, 55]
    assert np.abs(warped[0].numpy() - kat["kat5_warp_m1"]).max() < 1e-5


@pytest.mark.parametrize("s,H,W,automask", [(0, 24, 40, True), (1, 36, 72, True), (2, 32, 64, False), (0, 17, 33, True)])
def test_photometric_forward_backward_vs_oracle(s, H, W, automask):
    target, sources, disp, K, invK, Ts = _photo_case(H=H, W=W, s=s)
    B = target.shape[0]
    g = torch.Generator().manual_seed(5)
    noise = [1e-5 * torch.randn(B, 1, H, W, generator=g) for _ in sources]
    d0 = disp.clone().requires_grad_(True)
    T0 = [T.clone().requires_grad_(True) for T in Ts]
    m, idx, warped = O.photometric_scale(d0, target, sources, T0, K, invK, automask=automask, noise=noise)
    (m / 4).backward()
    d1 = disp.clone().requires_grad_(True)
    T1 = [T.clone().requires_grad_(True) for T in Ts]
    loss, winner, idx1, warped1 = JF.photometric_loss(d1, target, sources, T1, K, invK, num_scales=4, automask=automask,
                                                      noise=[n[:, 0] for n in noise], debug_outputs=True)
    assert abs(loss.item() - m.item() / 4) <= 1e-5 * abs(m.item() / 4)
    assert (idx1 != idx).float().mean().item() < 2e-3
    for w0, w1 in zip(warped, warped1):
        # sampling coordinates are O(W) in fp32: a few ulps of coordinate error move a textured pixel by ~1e-4
        assert (w0 - w1).abs().max().item() < 5e-4 and (w0 - w1).abs().mean().item() < 2e-6
    loss.backward()
    gd0, gd1 = d0.grad, d1.grad
    assert (gd0 - gd1).abs().max().item() <= 2e-3 * gd0.abs().max().item() + 1e-9
    for a, b in zip(T0, T1):
        assert (a.grad - b.grad).abs().max().item() <= 2e-3 * a.grad.abs().max().item() + 1e-9
