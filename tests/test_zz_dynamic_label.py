"""``type='dynamic'`` / ``'Argo_dynamic'`` CGT scale label (reference net.py:311-402) against the oracle, whose
dynamic branch is pinned by the reference's own run (tests/golden/e2e_dynamic_1024.npz, tests/test_oracle_e2e.py).

``[gpu]`` calls libjpb200.so through the C ABI on cuda:0; ``[emu]`` runs the same kernel source compiled as host C++.
(Sorted after the training-path suites: mode 2 of ``jpb_scale_label`` was added after this round's GPU budget was spent.)"""
import pytest
import torch

from test_losses import D, _label_case, dev  # noqa: F401  (``dev`` is the emu / gpu fixture)
from oracle import port as O

from jperceiver_b200 import functional as JF


@pytest.mark.parametrize("split", ["odometry", "argo"])
def test_scale_label_dynamic_vs_oracle(dev, split):
    """``get_scale_label_dynamic`` (net.py:311-402): z-map without the 0.27 m KITTI offset, masked by the cv2 quad only; the
    BEV label contributes its shape and nothing else (the oracle is pinned by tests/golden/e2e_dynamic_1024.npz)."""
    opt, inp, hw = _label_case(split)
    opt["type"] = "dynamic" if split == "odometry" else "Argo_dynamic"
    ref = O.scale_label(opt, inp)
    occ = opt["occ_map_size"]
    Minv = O.bev_to_image_homography(inp[("odometry_K", 0, 0)][:, :3, :3], inp[("Tr_cam2_velo", 0, 0)], split, occ)
    quad = (O.static_quad_mask(Minv[0], occ, *hw) > 0).to(torch.uint8)
    assert (ref > 0).sum().item() > 50
    for label in (D(inp[("bothS", 0, 0)], dev), None):
        got = JF.scale_label(label, D(inp[("odometry_K", 0, 0)], dev), D(inp[("Tr_cam2_velo", 0, 0)], dev), hw, split=split,
                             mode="dynamic", quad=D(quad, dev), occ=occ).cpu()
        assert ((got > 0) != (ref > 0)).float().mean().item() < 1e-3
        both = (got > 0) & (ref > 0)
        assert (got - ref)[both].abs().max().item() < 2e-3
    if split == "odometry":   # the static label of the same inputs differs by the 0.27 m offset (net.py:229-233 vs 323-326)
        opt["type"] = "static"
        assert not torch.equal(O.scale_label(opt, inp), ref)


def test_scale_loss_of_an_empty_label_is_nan_like_the_reference(dev):
    """``get_scale_loss`` takes ``torch.mean`` of a masked selection (net.py:207-210): with no labelled pixel the reference's loss is
    NaN (mean of an empty tensor) — the kernels reproduce that instead of inventing a zero."""
    g = torch.Generator().manual_seed(0)
    disp = torch.rand(2, 1, 16, 32, generator=g) * 0.9 + 0.05
    label = torch.zeros(2, 1, 40, 90)
    ref = O.scale_term(disp, label, "static")
    got = JF.scale_loss(D(disp, dev), D(label, dev), 0.1, False)
    assert ref.item() != ref.item() and got.item() != got.item()
