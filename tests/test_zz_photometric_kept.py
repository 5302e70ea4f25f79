"""Opt-in backward schedule of the fused photometric kernel that stages the forward's warped frames.

``[gpu]`` calls libjpb200.so through the C ABI on cuda:0; ``[emu]`` runs the same kernel source compiled as host C++.
(Sorted after the training-path suites: written after this round's GPU budget was spent; off by default.)"""
import pytest
import torch

from test_losses import D, _photo_case, dev  # noqa: F401  (``dev`` is the emu / gpu fixture)
from oracle import port as O

from jperceiver_b200 import functional as JF


@pytest.mark.parametrize("s,H,W,automask", [(0, 24, 40, True), (1, 36, 72, True), (0, 17, 33, False)])
def test_photometric_backward_from_kept_warped_frames(dev, s, H, W, automask):
    """Opt-in schedule (``keep_warped`` / JPB_PHOTO_KEEP_WARPED=1): the forward keeps outputs[("color",f,s)] and the backward stages
    them instead of re-projecting its 2-pixel apron.  Same loss; gradients equal to those of the default schedule (the staged
    values are the ones the forward computed) and to the oracle's."""
    target, sources, disp, K, invK, Ts = _photo_case(H=H, W=W, s=s, seed=11)
    grads = []
    for keep in (False, True):
        d = D(disp.detach().clone(), dev).requires_grad_(True)
        T = [D(t.detach().clone(), dev).requires_grad_(True) for t in Ts]
        loss = JF.photometric_loss(d, D(target, dev), D(sources, dev), T, D(K, dev), D(invK, dev), automask=automask, noise_scale=0.0,
                                   keep_warped=keep)[0]
        loss.backward()
        grads.append((loss.item(), d.grad.cpu(), [t.grad.cpu() for t in T]))
    (l0, gd0, gT0), (l1, gd1, gT1) = grads
    assert abs(l0 - l1) <= 1e-6 * abs(l0)
    scale = gd0.abs().max().item()
    # fp32 atomics' arrival order, and on the GPU the forward's fast reciprocal in the kept frames vs the exact division of the
    # re-projection: 1e-4 of the largest entry
    assert scale > 0 and (gd0 - gd1).abs().max().item() <= 1e-4 * scale
    for a, b in zip(gT0, gT1):
        assert (a - b).abs().max().item() <= 1e-4 * max(a.abs().max().item(), 1e-12)
    # and against the oracle's autograd
    d0 = disp.detach().clone().requires_grad_(True)
    T0 = [t.detach().clone().requires_grad_(True) for t in Ts]
    m, _, _ = O.photometric_scale(d0, target, sources, T0, K, invK, automask=automask, noise=None)
    (m / 4).backward()
    assert abs(l1 - (m / 4).item()) <= 1e-5 * abs(m.item())
    assert (gd1 - d0.grad).abs().max().item() <= 2e-3 * d0.grad.abs().max().item()
