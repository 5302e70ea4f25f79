"""The oracle port against the known-answer vectors produced by the reference's own functions
(tests/golden/kat.npz, written by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pat
from oracle import port as O


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLDEN, "kat.npz"))


def close(a, b, tol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


def test_kat0_disp_to_depth(kat):
    close(O.disp_to_depth(torch.tensor(0.5)).item(), kat["kat0"])
    assert abs(float(kat["kat0"]) - 0.199800193) < 1e-8  # the survey's printed value


def test_kat1_reprojection_error(kat):
    r = O.reprojection_error(pat((1, 3, 8, 12), 1), pat((1, 3, 8, 12), 2))
    close(r.numpy(), kat["kat1"])
    assert abs(r.mean().item() - 0.236392975) < 1e-7


def test_kat2_pose_matrix(kat):
    aa, t = torch.tensor([[.01, -.02, .03]]), torch.tensor([[.1, -.05, .2]])
    close(O.pose_matrix(aa, t, False).numpy(), kat["kat2_fwd"])
    close(O.pose_matrix(aa, t, True).numpy(), kat["kat2_inv"])


def _kat3_setup():
    H, W = 8, 12
    K = torch.tensor([[.58 * W, 0, .5 * W, 0], [0, 1.92 * H, .5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).unsqueeze(0)
    return K, torch.linalg.pinv(K)


def test_kat3_reprojection_grid(kat):
    K, invK = _kat3_setup()
    T = torch.from_numpy(kat["kat2_fwd"])
    depth = 1 + 4 * pat((1, 1, 8, 12), 3)
    grid = O.reproject_grid(depth, K, invK, T)
    close(grid.numpy(), kat["kat3_pix"], 2e-6)
    s = torch.nn.functional.grid_sample(pat((1, 3, 8, 12), 4), grid, padding_mode="border", align_corners=False)
    close(s.numpy(), kat["kat3_sample"], 2e-6)


def test_kat4_smoothness(kat):
    v = O.smooth_term(pat((1, 1, 4, 6), 5), pat((1, 3, 8, 12), 2), disp_norm=True)
    close(v.item(), kat["kat4"])
    assert abs(v.item() - 3.043148279) < 1e-6


def test_kat5_photometric_scale(kat):
    K, invK = _kat3_setup()
    Tf, Ti = torch.from_numpy(kat["kat2_fwd"]), torch.from_numpy(kat["kat2_inv"])
    m, idx, warped = O.photometric_scale(0.1 + 0.8 * pat((1, 1, 4, 6), 8), pat((1, 3, 8, 12), 2),
                                         [pat((1, 3, 8, 12), 6), pat((1, 3, 8, 12), 7)], [Ti, Tf], K, invK)
    close(m.item(), kat["kat5_mean"])
    assert np.array_equal(np.bincount(idx.flatten().numpy(), minlength=4), kat["kat5_hist"])
    assert list(kat["kat5_hist"]) == [0, 0, 41, 55]
    close(warped[0].numpy(), kat["kat5_warp_m1"], 2e-6)


def test_kat6_bev_terms(kat):
    logits = torch.cat([4 * pat((1, 1, 16, 16), 9) - 2, 4 * pat((1, 1, 16, 16), 10) - 2], 1)
    lab = torch.zeros(1, 1, 16, 16)
    lab[:, :, 4:11, 3:9] = 1
    close(O.signed_distance(lab[0, 0].numpy()), kat["kat6_sdf"], 1e-12)
    tot = O.bev_head_loss(logits, lab, 5.0, 20.0, 20.0).item()
    close(tot, 20 * kat["kat6_iou"] + kat["kat6_ce"] + 20 * kat["kat6_bd"], 1e-6)
    assert abs(float(kat["kat6_sdf"].sum()) - 761.522685) < 1e-5


def test_kat7_topview_loss(kat):
    big = torch.cat([4 * pat((2, 1, 256, 256), 9) - 2, 4 * pat((2, 1, 256, 256), 10) - 2], 1)
    lab = torch.zeros(2, 1, 256, 256)
    lab[:, :, 64:176, 48:144] = 1
    lab[1, :, 200:240, 10:250] = 1
    close(O.bev_head_loss(big[:1], lab[:1], 5.0).item(), kat["kat7_b1_w5"], 1e-6)
    close(O.bev_head_loss(big[:1], lab[:1], 15.0).item(), kat["kat7_b1_w15"], 1e-6)
    close(O.bev_head_loss(big, lab, 5.0).item(), kat["kat7_b2_w5"], 1e-6)


def test_kat8_transform_loss(kat):
    v = (pat((1, 128, 8, 8), 11) - pat((1, 128, 8, 8), 12)).abs().mean().item()
    close(v, kat["kat8"])


def test_empty_foreground_sdf_is_zero():
    assert not O.signed_distance(np.zeros((8, 8), dtype=np.uint8)).any()


@pytest.mark.parametrize("loss_type", ["iou", "dice", "tversky", "focal"])
@pytest.mark.parametrize("loss_sum", [1, 2, 3])
def test_bev_loss_variants_match_reference(loss_type, loss_sum):
    """every region-loss variant x loss_sum of compute_topview_loss against values produced by the
    reference's own code (oracle/make_golden_bev.py), value and input-gradient checksum."""
    gold = np.load(os.path.join(GOLDEN, "kat_bev_variants.npz"))
    big = torch.cat([4 * pat((2, 1, 256, 256), 9) - 2, 4 * pat((2, 1, 256, 256), 10) - 2], 1)
    lab = torch.zeros(2, 1, 256, 256)
    lab[:, :, 64:176, 48:144] = 1
    lab[1, :, 200:240, 10:250] = 1
    for w in (5, 15):
        x = big.clone().requires_grad_(True)
        v = O.bev_head_loss(x, lab, float(w), 20.0, 20.0, loss_type=loss_type, loss_sum=loss_sum)
        (g,) = torch.autograd.grad(v, x)
        key = "%s_s%d_w%d" % (loss_type, loss_sum, w)
        assert abs(v.item() - float(gold[key])) <= 2e-6 * max(abs(float(gold[key])), 1.0), key
        assert abs(g.abs().sum().item() - float(gold[key + "_gsum"])) <= 1e-4 * float(gold[key + "_gsum"]), key
