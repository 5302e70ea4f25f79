"""GPU parity of the PACKED forward schedule of the fused photometric kernel (``jpb_photometric_set_variant(3)``: the two
source frames of a snippet evaluated with one FADD2 / FMUL2 / FFMA2 per operation, csrc/photometric.cu namespace v3).  The
schedule is opt-in: it was written after this round's GPU budget was spent, so its logic is covered by the host emulation
(tests/test_losses.py, ``fwd3`` cases) and these are its first on-device checks — kept in a file that sorts last so that the
measured default path is always exercised first.  Same oracle, same tolerances as tests/test_losses.py."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pat

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import port as O  # noqa: E402
from test_losses import D, _photo_case  # noqa: E402

from jperceiver_b200 import _lib, functional as JF  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture()
def packed():
    _lib._handle, _lib._emulated = None, False
    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    _lib.check(_lib.lib().jpb_photometric_set_variant(3), "jpb_photometric_set_variant")
    yield torch.device("cuda:0")
    _lib.check(_lib.lib().jpb_photometric_set_variant(2), "jpb_photometric_set_variant")


def test_packed_kat5(packed):
    dev = packed
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    H, W = 8, 12
    K = torch.tensor([[.58 * W, 0, .5 * W, 0], [0, 1.92 * H, .5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).unsqueeze(0)
    Tf, Ti = torch.from_numpy(kat["kat2_fwd"]), torch.from_numpy(kat["kat2_inv"])
    loss, winner, idx, warped = JF.photometric_loss(
        D(0.1 + 0.8 * pat((1, 1, 4, 6), 8), dev), D(pat((1, 3, 8, 12), 2), dev),
        D([pat((1, 3, 8, 12), 6), pat((1, 3, 8, 12), 7)], dev), D([Ti, Tf], dev), D(K, dev), D(torch.linalg.pinv(K), dev),
        num_scales=1, noise_scale=0.0, debug_outputs=True)
    assert abs(loss.item() - float(kat["kat5_mean"])) < 2e-6
    assert list(np.bincount(idx.flatten().cpu().numpy(), minlength=4)) == [0, 0, 41, 55]
    assert np.abs(warped[0].cpu().numpy() - kat["kat5_warp_m1"]).max() < 1e-5


@pytest.mark.parametrize("s,H,W,automask,F", [(0, 24, 40, True, 2), (1, 36, 72, True, 2), (2, 32, 64, False, 2), (0, 17, 33, True, 2),
                                              (0, 40, 72, True, 1), (3, 64, 96, True, 2)])
def test_packed_forward_vs_oracle_and_backward_from_its_winner(packed, s, H, W, automask, F):
    dev = packed
    target, sources, disp, K, invK, Ts = _photo_case(H=H, W=W, s=s, F=F)
    B = target.shape[0]
    g = torch.Generator().manual_seed(5)
    noise = [1e-5 * torch.randn(B, 1, H, W, generator=g) for _ in sources]
    d0 = disp.clone().requires_grad_(True)
    T0 = [T.clone().requires_grad_(True) for T in Ts]
    m, idx, warped = O.photometric_scale(d0, target, sources, T0, K, invK, automask=automask, noise=noise)
    (m / 4).backward()
    d1 = D(disp, dev).requires_grad_(True)
    T1 = [D(T, dev).requires_grad_(True) for T in Ts]
    loss, winner, idx1, warped1 = JF.photometric_loss(d1, D(target, dev), D(sources, dev), T1, D(K, dev), D(invK, dev),
                                                      num_scales=4, automask=automask,
                                                      noise=D([n[:, 0] for n in noise], dev), debug_outputs=True)
    assert abs(loss.item() - m.item() / 4) <= 1e-5 * abs(m.item() / 4)          # tolerance: 1e-5 relative (north star: 1e-3)
    assert (idx1.cpu() != idx).float().mean().item() < 2e-3
    assert (winner.cpu().long() != idx1.cpu()).sum().item() == 0
    for w0, w1 in zip(warped, warped1):
        assert (w0 - w1.cpu()).abs().max().item() < 5e-4 and (w0 - w1.cpu()).abs().mean().item() < 2e-6
    loss.backward()                                                             # the backward kernel consumes this forward's arg-min
    gd0, gd1 = d0.grad, d1.grad.cpu()
    assert (gd0 - gd1).abs().max().item() <= 2e-3 * gd0.abs().max().item() + 1e-9
    for a, b in zip(T0, T1):
        assert (a.grad - b.grad.cpu()).abs().max().item() <= 2e-3 * a.grad.abs().max().item() + 1e-9


def test_packed_full_size_vs_default_schedule_and_oracle(packed):
    """BASELINE size (320x1024, B=4, F=2), every scale: the packed schedule against the default one and (scale 0) the oracle;
    Philox noise deterministic per (seed, stream)."""
    dev = packed
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from bench_photometric import make_case
    B, H, W, F = 4, 320, 1024, 2
    target, sources, disps, K, invK, Ts = make_case(B, H, W, F, dev)
    for s in range(4):
        l3, w3, _, _ = JF.photometric_loss(disps[s], target, sources, Ts, K, invK, num_scales=4, noise_scale=0.0)
        _lib.check(_lib.lib().jpb_photometric_set_variant(2), "jpb_photometric_set_variant")
        l2, w2, _, _ = JF.photometric_loss(disps[s], target, sources, Ts, K, invK, num_scales=4, noise_scale=0.0)
        _lib.check(_lib.lib().jpb_photometric_set_variant(3), "jpb_photometric_set_variant")
        assert abs(l3.item() - l2.item()) <= 1e-5 * abs(l2.item()), s
        assert (w3 != w2).float().mean().item() < 1e-3, s
    m, idx, _ = O.photometric_scale(disps[0].cpu(), target.cpu(), [x.cpu() for x in sources], [T.cpu() for T in Ts],
                                    K.cpu(), invK.cpu(), automask=True, noise=None)
    l3, w3, _, _ = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, num_scales=4, noise_scale=0.0)
    assert abs(l3.item() - m.item() / 4) <= 1e-4 * abs(m.item() / 4)
    assert (w3.cpu().long() != idx).float().mean().item() < 1e-3
    la = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, seed=7, stream=1)[0].item()
    lb = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, seed=7, stream=1)[0].item()
    lc = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, seed=8, stream=1)[0].item()
    assert la == lb and abs(la - l3.item()) < 1e-4 and abs(lc - l3.item()) < 1e-4
