"""Parity of the fused loss-chain kernels with the oracle port (values and autograd gradients).

Every case runs twice: ``[gpu]`` (marked ``gpu``) calls libjpb200.so through the C ABI on cuda:0 —
the parity test proper; ``[emu]`` runs the same kernel sources compiled as host C++ (tests/emu), a
logic check that works in the GPU-less authoring container."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pat

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402
from emu.torch_ops import torch_conv  # noqa: E402
from oracle import port as O  # noqa: E402

from jperceiver_b200 import _lib, functional as JF  # noqa: E402


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


def D(x, dev):
    if isinstance(x, (list, tuple)):
        return [D(t, dev) for t in x]
    return x.to(dev)


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


@pytest.fixture(params=[2, 3], ids=["fwd2", "fwd3"])
def photo_variant(request, dev):
    """Both forward schedules of the fused photometric kernel (include/jpb200.h: jpb_photometric_set_variant).  The packed
    one is opt-in until it has been measured; its GPU parity cases live in tests/test_zzy_photometric_packed.py (sorted last)."""
    if request.param == 3 and dev.type == "cuda":
        pytest.skip("packed schedule on the GPU: tests/test_zzy_photometric_packed.py")
    _lib.check(_lib.lib().jpb_photometric_set_variant(request.param), "jpb_photometric_set_variant")
    yield request.param
    _lib.check(_lib.lib().jpb_photometric_set_variant(2), "jpb_photometric_set_variant")


def test_photometric_kat5(dev, photo_variant):
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


@pytest.mark.parametrize("s,H,W,automask", [(0, 24, 40, True), (1, 36, 72, True), (2, 32, 64, False), (0, 17, 33, True)])
def test_photometric_forward_backward_vs_oracle(dev, photo_variant, s, H, W, automask):
    target, sources, disp, K, invK, Ts = _photo_case(H=H, W=W, s=s)
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
    assert abs(loss.item() - m.item() / 4) <= 1e-5 * abs(m.item() / 4)
    assert (idx1.cpu() != idx).float().mean().item() < 2e-3
    assert (winner.cpu().long() != idx1.cpu()).sum().item() == 0
    for w0, w1 in zip(warped, warped1):
        # sampling coordinates are O(W) in fp32: a few ulps of coordinate error move a textured pixel by ~1e-4
        assert (w0 - w1.cpu()).abs().max().item() < 5e-4 and (w0 - w1.cpu()).abs().mean().item() < 2e-6
    loss.backward()
    gd0, gd1 = d0.grad, d1.grad.cpu()
    assert (gd0 - gd1).abs().max().item() <= 2e-3 * gd0.abs().max().item() + 1e-9
    for a, b in zip(T0, T1):
        assert (a.grad - b.grad.cpu()).abs().max().item() <= 2e-3 * a.grad.abs().max().item() + 1e-9


def test_photometric_single_source_and_inkernel_noise(dev, photo_variant):
    """frame_ids [0, -1] (configs 1 and 4: one source frame) and the in-kernel Philox draw of the automask noise."""
    target, sources, disp, K, invK, Ts = _photo_case(H=40, W=72, s=0, F=1, seed=3)
    m, idx, warped = O.photometric_scale(disp, target, sources, Ts, K, invK, automask=True, noise=None)
    args = (D(disp, dev), D(target, dev), D(sources, dev), D(Ts, dev), D(K, dev), D(invK, dev))
    loss, winner, idx1, warped1 = JF.photometric_loss(*args, num_scales=4, noise_scale=0.0, debug_outputs=True)
    assert abs(loss.item() - m.item() / 4) <= 1e-5 * abs(m.item() / 4)
    assert (idx1.cpu() != idx).float().mean().item() < 2e-3 and int(idx1.max()) <= 1
    assert (warped[0] - warped1[0].cpu()).abs().max().item() < 5e-4
    la = JF.photometric_loss(*args, num_scales=4, seed=7, stream=1)[0].item()
    lb = JF.photometric_loss(*args, num_scales=4, seed=7, stream=1)[0].item()
    lc = JF.photometric_loss(*args, num_scales=4, seed=8, stream=1)[0].item()
    assert la == lb and la != lc
    assert abs(la - loss.item()) < 1e-4 * abs(loss.item()) + 1e-6 and abs(lc - loss.item()) < 1e-4 * abs(loss.item()) + 1e-6


def test_photometric_variants_agree(dev):
    """The packed forward schedule reproduces the default one: same arg-min except at fp32 ties, same loss to 1e-6 relative."""
    if dev.type == "cuda":
        pytest.skip("packed schedule on the GPU: tests/test_zzy_photometric_packed.py")
    target, sources, disp, K, invK, Ts = _photo_case(B=2, H=64, W=96, s=1, seed=11)
    args = (D(disp, dev), D(target, dev), D(sources, dev), D(Ts, dev), D(K, dev), D(invK, dev))
    out = {}
    try:
        for v in (2, 3):
            _lib.check(_lib.lib().jpb_photometric_set_variant(v), "jpb_photometric_set_variant")
            out[v] = JF.photometric_loss(*args, num_scales=4, noise_scale=0.0, debug_outputs=True)
    finally:
        _lib.check(_lib.lib().jpb_photometric_set_variant(2), "jpb_photometric_set_variant")
    assert abs(out[2][0].item() - out[3][0].item()) <= 2e-6 * abs(out[2][0].item())
    assert (out[2][2] != out[3][2]).float().mean().item() < 2e-3
    for w2, w3 in zip(out[2][3], out[3][3]):
        assert (w2 - w3).abs().max().item() < 5e-4
    assert _lib.lib().jpb_photometric_set_variant(7) != 0        # unknown schedule: argument error, nothing changes


@pytest.mark.parametrize("s,H,W,F,keep", [(0, 24, 40, 2, False), (1, 36, 72, 2, True), (3, 64, 96, 2, True), (2, 48, 80, 1, False), (0, 17, 33, 2, True)])
def test_photometric_backward_schedules_agree(dev, s, H, W, F, keep):
    """jpb_photometric_set_bwd_variant: the register-resident kernel (4, default for F <= 2) against the generic one (1) — disparity
    and pose gradients equal within fp32 summation order, with re-projected and with kept warped frames, at every scale's
    up-sampling factor (the separable transposed up-sampling replaces the shared-memory atomics of schedule 1)."""
    target, sources, disp, K, invK, Ts = _photo_case(B=2, H=H, W=W, s=s, F=F, seed=5)
    grads = {}
    try:
        for v in (1, 4):
            _lib.check(_lib.lib().jpb_photometric_set_bwd_variant(v), "jpb_photometric_set_bwd_variant")
            d = D(disp.clone(), dev).requires_grad_(True)
            T = [D(t.clone(), dev).requires_grad_(True) for t in Ts]
            loss = JF.photometric_loss(d, D(target, dev), D(sources, dev), T, D(K, dev), D(invK, dev), num_scales=4, noise_scale=0.0,
                                       keep_warped=keep)[0]
            loss.backward()
            grads[v] = (d.grad.cpu(), [t.grad.cpu() for t in T])
    finally:
        _lib.check(_lib.lib().jpb_photometric_set_bwd_variant(4), "jpb_photometric_set_bwd_variant")
    scale = grads[1][0].abs().max().item()
    # same formulas, different FMA contraction of the cancelling window statistics on the GPU: 1e-4 of the largest entry
    assert scale > 0 and (grads[1][0] - grads[4][0]).abs().max().item() <= 1e-4 * scale
    for a, b in zip(grads[1][1], grads[4][1]):
        assert (a - b).abs().max().item() <= 1e-4 * a.abs().max().item() + 1e-12
    assert _lib.lib().jpb_photometric_set_bwd_variant(3) != 0      # unknown schedule: argument error


@pytest.mark.parametrize("F,explicit_noise", [(2, True), (1, False)])
def test_photometric_identity_terms_shared_across_scales(dev, F, explicit_noise):
    """JpbPhotoArgs.ident_mode: the identity candidates (target vs un-warped sources, net.py:159-166) do not depend on the
    scale, so the first launch of a step stores their errors and the other scales' launches read them.  Every scale's loss,
    arg-min and gradients equal those of the launches that evaluate the identity terms themselves."""
    H, W = 48, 80
    target, sources, _, K, invK, Ts = _photo_case(B=2, H=H, W=W, s=0, F=F, seed=9)
    g = torch.Generator().manual_seed(3)
    disps = [0.05 + 0.9 * torch.rand(2, 1, H >> (s + 1), W >> (s + 1), generator=g) for s in range(4)]
    noises = [[1e-5 * torch.randn(2, H, W, generator=g) for _ in sources] for _ in range(4)] if explicit_noise else [None] * 4
    _lib.check(_lib.lib().jpb_photometric_set_variant(3), "jpb_photometric_set_variant")
    try:
        res = {}
        for shared in (False, True):
            cache = {} if shared else None
            out = []
            for s in range(4):
                d = D(disps[s].clone(), dev).requires_grad_(True)
                loss, winner, idx, _ = JF.photometric_loss(d, D(target, dev), D(sources, dev), D(Ts, dev), D(K, dev), D(invK, dev), num_scales=4,
                                                           noise=D(noises[s], dev) if noises[s] is not None else None, noise_scale=0.0,
                                                           debug_outputs=True, ident_cache=cache)
                loss.backward()
                out.append((loss.item(), idx.cpu(), d.grad.cpu()))
            res[shared] = out
            if shared:
                assert cache["err"].shape == (2, H, W, 2)
    finally:
        _lib.check(_lib.lib().jpb_photometric_set_variant(3), "jpb_photometric_set_variant")     # the library default
    for (l0, i0, g0), (l1, i1, g1) in zip(res[False], res[True]):
        assert l0 == l1 and (i0 != i1).sum().item() == 0
        assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item()      # fp32 atomics: arrival order


def test_area_pyramid_and_smoothness_kat4(dev):
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    img = pat((1, 3, 8, 12), 2)
    J = JF.area_pyramid(D(img, dev), 1)[0]
    assert (J.cpu() - torch.nn.functional.interpolate(img, (4, 6), mode="area")).abs().max().item() < 1e-6
    v = JF.smooth_loss(D(pat((1, 1, 4, 6), 5), dev), J, 1.0, True)
    assert abs(v.item() - float(kat["kat4"])) < 1e-5


@pytest.mark.parametrize("disp_norm", [True, False])
def test_smoothness_forward_backward_vs_oracle(dev, disp_norm):
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 48, 80
    img = torch.rand(B, 3, H, W, generator=g)
    levels = JF.area_pyramid(D(img, dev), 4)
    for s in range(4):
        h, w = H >> (s + 1), W >> (s + 1)
        assert (levels[s].cpu() - torch.nn.functional.interpolate(img, (h, w), mode="area")).abs().max().item() < 1e-6
        disp = torch.rand(B, 1, h, w, generator=g) * 0.9 + 0.05
        d0 = disp.clone().requires_grad_(True)
        ref = 1e-3 * O.smooth_term(d0, img, disp_norm) / (2 ** s) / 4
        ref.backward()
        d1 = D(disp, dev).requires_grad_(True)
        got = JF.smooth_loss(d1, levels[s], 1e-3 / (2 ** s) / 4, disp_norm)
        assert abs(got.item() - ref.item()) <= 2e-5 * abs(ref.item())
        (got * 3.0).backward()
        assert (d1.grad.cpu() / 3.0 - d0.grad).abs().max().item() <= 1e-3 * d0.grad.abs().max().item()


def _label_case(split, B=2, occ=64, seed=0):
    from oracle.ref_loader import default_options
    opt = default_options(type="Argo_both", split=split, occ_map_size=occ, height=4 * occ, width=4 * occ)
    hw = (120, 400) if split != "argo" else (200, 240)
    inp = O.synth_inputs(opt, B, seed=seed, hw_full=hw)
    if split == "argo":  # 4x4 intrinsics as the Argoverse loader emits
        K4 = torch.eye(4).repeat(B, 1, 1)
        K4[:, :3, :3] = inp[("odometry_K", 0, 0)]
        inp[("odometry_K", 0, 0)] = K4
    # bring the projected BEV region inside the small test image
    inp[("odometry_K", 0, 0)][:, 0, 0] *= 0.3
    inp[("odometry_K", 0, 0)][:, 1, 1] *= 0.3
    inp[("odometry_K", 0, 0)][:, 0, 2] = hw[1] / 2
    inp[("odometry_K", 0, 0)][:, 1, 2] = hw[0] / 3
    return opt, inp, hw


@pytest.mark.parametrize("split,align", [("odometry", True), ("argo", True), ("odometry", False)])
def test_scale_label_both_vs_oracle(dev, split, align):
    opt, inp, hw = _label_case(split)
    ref = O.scale_label(opt, inp, warp_align_corners=align)
    got = JF.scale_label(D(inp[("both_dynamic", 0, 0)], dev), D(inp[("odometry_K", 0, 0)], dev),
                         D(inp[("Tr_cam2_velo", 0, 0)], dev), hw, split=split, mode="both", align_corners=align).cpu()
    assert (ref > 0).float().mean().item() > 0.01, "test geometry must put the BEV map inside the image"
    d = (got - ref).abs()
    # near the horizon the warp is ill-conditioned: the reference's own fp32 matrix inverses move a few pixels
    assert (d > 2e-3).float().mean().item() < 1e-3 and d.mean().item() < 2e-5 and d.max().item() < 0.05


def test_scale_label_static_and_loss_vs_oracle(dev):
    opt, inp, hw = _label_case("odometry")
    opt["type"] = "static"
    ref = O.scale_label(opt, inp)
    occ = opt["occ_map_size"]
    Minv = O.bev_to_image_homography(inp[("odometry_K", 0, 0)][:, :3, :3], inp[("Tr_cam2_velo", 0, 0)], "odometry", occ)
    quad = (O.static_quad_mask(Minv[0], occ, *hw) > 0).to(torch.uint8)
    got = JF.scale_label(D(inp[("bothS", 0, 0)], dev), D(inp[("odometry_K", 0, 0)], dev), D(inp[("Tr_cam2_velo", 0, 0)], dev),
                         hw, split="odometry", mode="static", quad=D(quad, dev)).cpu()
    assert (ref > 0).sum().item() > 50
    assert ((got > 0) != (ref > 0)).float().mean().item() < 1e-3
    both = (got > 0) & (ref > 0)
    assert (got - ref)[both].abs().max().item() < 2e-3
    g = torch.Generator().manual_seed(1)
    for s, crop in ((0, False), (2, False)):
        disp = torch.rand(2, 1, 32 >> s, 64 >> s, generator=g) * 0.9 + 0.05
        d0 = disp.clone().requires_grad_(True)
        r = 0.1 * O.scale_term(d0, ref, "static") / (2 ** s) / 4
        r.backward()
        d1 = D(disp, dev).requires_grad_(True)
        v = JF.scale_loss(d1, D(ref, dev), 0.1 / (2 ** s) / 4, crop)
        assert abs(v.item() - r.item()) <= 1e-5 * abs(r.item())
        v.backward()
        assert (d1.grad.cpu() - d0.grad).abs().max().item() <= 1e-4 * d0.grad.abs().max().item()


def test_signed_distance_exact(dev):
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    lab = torch.zeros(1, 16, 16)
    lab[:, 4:11, 3:9] = 1
    assert np.abs(JF.signed_distance(D(lab, dev))[0].cpu().numpy() - kat["kat6_sdf"]).max() < 1e-6
    g = torch.Generator().manual_seed(7)
    maps = (torch.rand(5, 48, 48, generator=g) > 0.7).float()
    maps[3] = 0            # empty foreground -> all zeros
    maps[4, 10:30, 5:40] = 1
    got = JF.signed_distance(D(maps, dev)).cpu().numpy()
    for b in range(5):
        assert np.abs(got[b] - O.signed_distance(maps[b].numpy())).max() < 1e-5
    assert not got[3].any()


@pytest.mark.parametrize("channels_last", [False, True])
def test_bev_head_loss_forward_backward_vs_oracle(dev, channels_last):
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    big = torch.cat([4 * pat((2, 1, 256, 256), 9) - 2, 4 * pat((2, 1, 256, 256), 10) - 2], 1)
    lab = torch.zeros(2, 1, 256, 256)
    lab[:, :, 64:176, 48:144] = 1
    lab[1, :, 200:240, 10:250] = 1
    sdf = JF.signed_distance(D(lab[:, 0], dev))
    v = JF.bev_head_loss(D(big, dev), D(lab, dev), sdf, 5.0, 20.0, 20.0)
    assert abs(v.item() - float(kat["kat7_b2_w5"])) <= 2e-6 * abs(float(kat["kat7_b2_w5"]))
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(3, 2, 32, 32, generator=g) * 2
    labels = (torch.rand(3, 1, 32, 32, generator=g) > 0.6).float()
    labels[2] = 0
    l0 = logits.clone().requires_grad_(True)
    ref = O.bev_head_loss(l0, labels, 15.0, 20.0, 20.0)
    ref.backward()
    l1 = D(logits, dev)
    if channels_last:
        l1 = l1.contiguous(memory_format=torch.channels_last)
    l1.requires_grad_(True)
    got = JF.bev_head_loss(l1, D(labels, dev), JF.signed_distance(D(labels[:, 0], dev)), 15.0, 20.0, 20.0)
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())
    (2.0 * got).backward()
    assert (l1.grad.cpu() / 2.0 - l0.grad).abs().max().item() <= 1e-4 * l0.grad.abs().max().item()


def test_l1_mean_vs_kat8(dev):
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    x = D(pat((1, 128, 8, 8), 11), dev).requires_grad_(True)
    y = D(pat((1, 128, 8, 8), 12), dev).requires_grad_(True)
    v = JF.l1_mean(x, y)
    assert abs(v.item() - float(kat["kat8"])) < 1e-6
    v.backward()
    ref = torch.sign(x.detach() - y.detach()) / x.numel()
    assert x.grad.device == x.device
    assert (x.grad - ref).abs().max().item() < 1e-9 and (y.grad + ref).abs().max().item() < 1e-9


@pytest.mark.gpu
def test_photometric_full_size_vs_oracle_and_properties():
    """BASELINE size (320x1024, B=4, F=2), scale 0: oracle parity, plus size-independent properties:
    batch-permutation invariance of the mean, and linearity of the gradient in the upstream scalar."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from bench_photometric import make_case
    _lib._handle, _lib._emulated = None, False
    dev = torch.device("cuda:0")
    B, H, W, F = 4, 320, 1024, 2
    target, sources, disps, K, invK, Ts = make_case(B, H, W, F, dev)
    disp = disps[0].clone().requires_grad_(True)
    loss, winner, _, _ = JF.photometric_loss(disp, target, sources, Ts, K, invK, num_scales=4, noise_scale=0.0)
    m, idx, _ = O.photometric_scale(disps[0].cpu(), target.cpu(), [s.cpu() for s in sources], [T.cpu() for T in Ts],
                                    K.cpu(), invK.cpu(), automask=True, noise=None)
    assert abs(loss.item() - m.item() / 4) <= 1e-4 * abs(m.item() / 4)          # tolerance: 1e-4 relative (north star: 1e-3)
    assert (winner.cpu().long() != idx).float().mean().item() < 1e-3
    perm = torch.tensor([2, 0, 3, 1], device=dev)
    loss_p, _, _, _ = JF.photometric_loss(disps[0][perm], target[perm], [s[perm] for s in sources], [T[perm] for T in Ts],
                                          K[perm], invK[perm], num_scales=4, noise_scale=0.0)
    assert abs(loss_p.item() - loss.item()) <= 1e-6 * abs(loss.item())
    (g1,) = torch.autograd.grad(loss, disp, retain_graph=True)
    (g3,) = torch.autograd.grad(3.0 * loss, disp)
    assert (g3 - 3.0 * g1).abs().max().item() <= 1e-5 * g1.abs().max().item() * 3
    # in-kernel Philox noise: deterministic per (seed, stream), different across seeds, ~1e-5 effect
    la = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, seed=7, stream=1)[0].item()
    lb = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, seed=7, stream=1)[0].item()
    lc = JF.photometric_loss(disps[0], target, sources, Ts, K, invK, seed=8, stream=1)[0].item()
    assert la == lb and abs(la - loss.item()) < 1e-4 and abs(lc - loss.item()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_photometric_backward_schedules_agree_full_size(s):
    """BASELINE frame size (320x1024, F=2), every scale, with the warped frames the forward kept (the drop-in default): both
    backward schedules against the oracle's autograd evaluated in float64.  The window statistics cancel (variance = E[x^2] -
    E[x]^2 of smooth frames), so two fp32 evaluations with different FMA contraction differ by ~1e-3 of the largest gradient and
    the fp32 oracle itself is 2e-2 away from the float64 one (arg-min flips at ties); the default schedule has to be as close to
    float64 as the generic one, in the L2 norm, and the two have to agree entry-wise within that fp32 noise."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from bench_photometric import make_case
    _lib._handle, _lib._emulated = None, False
    dev = torch.device("cuda:0")
    target, sources, disps, K, invK, Ts = make_case(2, 320, 1024, 2, dev)
    dd = torch.float64
    d64 = disps[s].detach().cpu().to(dd).requires_grad_(True)
    T64 = [t.detach().cpu().to(dd).requires_grad_(True) for t in Ts]
    m, _, _ = O.photometric_scale(d64, target.cpu().to(dd), [x.cpu().to(dd) for x in sources], T64, K.cpu().to(dd), invK.cpu().to(dd),
                                  automask=True, noise=None)
    (m / 4).backward()
    grads = {}
    try:
        for v in (1, 4):
            _lib.check(_lib.lib().jpb_photometric_set_bwd_variant(v), "jpb_photometric_set_bwd_variant")
            d = disps[s].detach().clone().requires_grad_(True)
            T = [t.detach().clone().requires_grad_(True) for t in Ts]
            JF.photometric_loss(d, target, sources, T, K, invK, num_scales=4, noise_scale=0.0, debug_outputs=True)[0].backward()
            grads[v] = (d.grad.cpu().to(dd), [t.grad.cpu().to(dd) for t in T])
    finally:
        _lib.check(_lib.lib().jpb_photometric_set_bwd_variant(4), "jpb_photometric_set_bwd_variant")
    ref = d64.grad
    e1 = ((grads[1][0] - ref).norm() / ref.norm()).item()
    e4 = ((grads[4][0] - ref).norm() / ref.norm()).item()
    assert e4 <= max(1.25 * e1, 2e-3), (e1, e4)
    assert e4 < 5e-2, (e1, e4)
    assert (grads[1][0] - grads[4][0]).abs().max().item() <= 2e-2 * ref.abs().max().item()
    for a, b, r in zip(grads[1][1], grads[4][1], T64):
        assert (b - r.grad).abs().max().item() <= max(1.25 * (a - r.grad).abs().max().item(), 2e-3 * r.grad.abs().max().item())


@pytest.mark.parametrize("k,s,p,H,W,C", [(5, 1, 2, 12, 20, 8), (5, 1, 2, 45, 9, 4), (3, 2, 1, 16, 24, 16), (2, 2, 0, 8, 8, 128), (3, 2, 1, 15, 21, 4)])
def test_maxpool_forward_backward_vs_torch(dev, k, s, p, H, W, C):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, C, H, W, generator=g)
    x[0, :, 3:6, 3:6] = 1.5      # plateaus: ties must resolve to the first maximum, like ATen
    x0 = x.clone().requires_grad_(True)
    ref = torch.nn.functional.max_pool2d(x0, k, s, p)
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    x1 = D(x, dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = JF.maxpool(x1, k, s, p)
    assert torch.equal(got.cpu(), ref.detach())
    got.backward(D(gy, dev))
    # overlapping windows accumulate by atomics: summation order differs from ATen's
    assert (x1.grad.cpu() - x0.grad).abs().max().item() <= 1e-5 * max(1.0, x0.grad.abs().max().item())


@pytest.mark.parametrize("C,act,want_bias", [(256, "leaky", True), (64, "relu", True), (1028, "sigmoid", True), (34, "leaky", True), (6, "relu", True),
                                             (128, "none", True), (96, "leaky", False)])
def test_act_bwd_vs_torch(dev, C, act, want_bias):
    """jpb_act_bwd (backward of the convolution epilogue: dz = dy * act'(y) and the bias gradient) in its 16-byte form (C % 4 == 0,
    C >= 32, incl. more than 256 channel groups), the scalar form (C = 34) and the narrow form (C = 6) against torch."""
    from jperceiver_b200 import conv as JC
    g = torch.Generator().manual_seed(C)
    B, H, W = 2, 7, 9
    z = torch.randn(B, C, H, W, generator=g)
    y = {"leaky": torch.nn.functional.leaky_relu(z, 0.01), "relu": torch.relu(z), "sigmoid": torch.sigmoid(z), "none": z}[act]
    gy = torch.randn(B, C, H, W, generator=g)
    z0 = z.clone().requires_grad_(True)
    y0 = {"leaky": torch.nn.functional.leaky_relu(z0, 0.01), "relu": torch.relu(z0), "sigmoid": torch.sigmoid(z0), "none": z0 * 1.0}[act]
    y0.backward(gy)
    cl = torch.channels_last
    dz, gb = JC.act_bwd(D(gy, dev).contiguous(memory_format=cl), D(y, dev).contiguous(memory_format=cl), act, want_bias)
    assert (dz.cpu() - z0.grad).abs().max().item() <= 1e-6 * max(1.0, z0.grad.abs().max().item())
    if want_bias:
        ref = z0.grad.sum((0, 2, 3))
        assert (gb.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    else:
        assert gb is None


def test_sum_n_equals_the_chain_of_binary_adds(dev):
    """jpb_sum_n (the residual chain of a CRP block, layers.py:186-199) is bit-identical to x + t1 + t2 + t3 + t4 evaluated left
    to right, and passes the upstream gradient to every term."""
    g = torch.Generator().manual_seed(12)
    xs = [D(torch.randn(2, 8, 5, 6, generator=g) * 10 ** (k - 2), dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) for k in range(5)]
    y = JF.sum_n(xs)
    ref = xs[0].detach()
    for t in xs[1:]:
        ref = t.detach() + ref
    assert torch.equal(y.detach().cpu(), ref.cpu())
    gy = D(torch.randn(2, 8, 5, 6, generator=g), dev)
    y.backward(gy)
    for t in xs:
        assert torch.equal(t.grad.cpu(), gy.cpu())


@pytest.mark.parametrize("variant", [1, 2])
def test_maxpool_backward_schedules(dev, variant):
    """jpb_maxpool_set_bwd_variant: scatter for every overlapping window (1) and the deterministic 5x5 gather (2) give the
    gradients of the default schedule (5x5 scatter, 3x3 / stride 2 gather)."""
    g = torch.Generator().manual_seed(8)
    res = {}
    try:
        for v in (0, variant):
            _lib.check(_lib.lib().jpb_maxpool_set_bwd_variant(v), "jpb_maxpool_set_bwd_variant")
            out = []
            for k, s, p, H, W, C in [(5, 1, 2, 23, 17, 8), (3, 2, 1, 16, 24, 16)]:
                gg = torch.Generator().manual_seed(k)
                x = D(torch.randn(2, C, H, W, generator=gg), dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
                y = JF.maxpool(x, k, s, p)
                y.backward(D(torch.randn(y.shape, generator=gg), dev))
                out.append(x.grad.cpu())
            res[v] = out
    finally:
        _lib.check(_lib.lib().jpb_maxpool_set_bwd_variant(0), "jpb_maxpool_set_bwd_variant")
    for a, b in zip(res[0], res[variant]):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, a.abs().max().item())
    assert _lib.lib().jpb_maxpool_set_bwd_variant(7) != 0


@pytest.mark.parametrize("C,N,up,reflect,act", [(256, 1, 1, 1, "sigmoid"), (16, 2, 0, 1, "none"), (32, 2, 0, 0, "leaky"), (64, 1, 1, 0, "none")])
def test_small_n_conv_forward_backward_vs_torch(dev, C, N, up, reflect, act):
    """The CUDA-core project / shift-and-add kernels for the 1-/2-channel heads (disparity, topview) against the
    library convolution: forward, weight gradient and data gradient."""
    from jperceiver_b200 import conv as JC
    g = torch.Generator().manual_seed(6)
    B, Hs, Ws = 2, 6, 10
    x = torch.randn(B, C, Hs, Ws, generator=g)
    w = torch.randn(N, C, 3, 3, generator=g) / (C * 9) ** 0.5
    b = torch.randn(N, generator=g)
    x0, w0 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = torch_conv([x0], [bool(up)], w0, b, 1, 1, bool(reflect), act, None)
    x1 = D(x, dev).contiguous(memory_format=torch.channels_last)
    w1 = D(w, dev).contiguous(memory_format=torch.channels_last)
    got = JC.smalln_fwd(x1, bool(up), w1, D(b, dev), bool(reflect), act)
    assert (got.cpu() - ref.detach()).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    z = torch_conv([x0], [bool(up)], w0, b, 1, 1, bool(reflect), "none", None)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    gw, gx = JC.smalln_bwd(x1, bool(up), D(dz, dev).contiguous(memory_format=torch.channels_last), w1, bool(reflect))
    assert (gw.cpu() - w0.grad).abs().max().item() <= 2e-5 * max(1.0, w0.grad.abs().max().item())
    assert (gx.cpu() - x0.grad).abs().max().item() <= 2e-5 * max(1.0, x0.grad.abs().max().item())


@pytest.mark.parametrize("C,H,W,relu,has_res", [(64, 12, 20, True, True), (16, 32, 32, True, False), (512, 3, 5, False, False), (128, 8, 8, False, True)])
def test_batchnorm_train_forward_backward_vs_torch(dev, C, H, W, relu, has_res):
    """Fused BN(+residual)(+ReLU): output, running statistics and all gradients against nn.functional.batch_norm."""
    g = torch.Generator().manual_seed(8)
    B = 3
    x = torch.randn(B, C, H, W, generator=g) * 2 + 0.5
    res = torch.randn(B, C, H, W, generator=g) if has_res else None
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    x0, g0, b0 = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    r0 = res.clone().requires_grad_(True) if has_res else None
    rm0, rv0 = rm.clone(), rv.clone()
    ref = torch.nn.functional.batch_norm(x0, rm0, rv0, g0, b0, True, 0.19, 1e-5)
    if has_res:
        ref = ref + r0
    if relu:
        ref = torch.relu(ref)
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    cl = torch.channels_last
    x1 = D(x, dev).contiguous(memory_format=cl).requires_grad_(True)
    g1, b1 = D(gamma, dev).requires_grad_(True), D(beta, dev).requires_grad_(True)
    r1 = D(res, dev).contiguous(memory_format=cl).requires_grad_(True) if has_res else None
    rm1, rv1 = D(rm, dev), D(rv, dev)
    got = JF.batchnorm_train(x1, r1, g1, b1, rm1, rv1, 0.19, 1e-5, relu)
    assert (got.cpu() - ref.detach()).abs().max().item() < 2e-5
    assert (rm1.cpu() - rm0).abs().max().item() < 1e-6 and (rv1.cpu() - rv0).abs().max().item() < 2e-6
    got.backward(D(gy, dev))
    assert (x1.grad.cpu() - x0.grad).abs().max().item() <= 2e-5 * max(1.0, x0.grad.abs().max().item())
    assert (g1.grad.cpu() - g0.grad).abs().max().item() <= 2e-5 * max(1.0, g0.grad.abs().max().item())
    assert (b1.grad.cpu() - b0.grad).abs().max().item() <= 2e-5 * max(1.0, b0.grad.abs().max().item())
    if has_res:
        assert (r1.grad.cpu() - r0.grad).abs().max().item() < 1e-6
    ev = JF.batchnorm_eval(x1.detach(), r1.detach() if has_res else None, g1.detach(), b1.detach(), rm1, rv1, 1e-5, relu)
    rev = torch.nn.functional.batch_norm(x, rm0, rv0, gamma, beta, False, 0.1, 1e-5)
    rev = rev + res if has_res else rev
    rev = torch.relu(rev) if relu else rev
    assert (ev.cpu() - rev).abs().max().item() < 2e-5


@pytest.mark.parametrize("loss_type", ["iou", "dice", "tversky", "focal"])
@pytest.mark.parametrize("loss_sum", [1, 2, 3])
def test_bev_loss_variants_vs_oracle_and_reference(dev, loss_type, loss_sum):
    """the region-loss variants of compute_topview_loss.  Values against the vectors produced by the
    reference's own code (tests/golden/kat_bev_variants.npz), gradients against the oracle's autograd."""
    gold = np.load(os.path.join(GOLDEN, "kat_bev_variants.npz"))
    big = torch.cat([4 * pat((2, 1, 256, 256), 9) - 2, 4 * pat((2, 1, 256, 256), 10) - 2], 1)
    lab = torch.zeros(2, 1, 256, 256)
    lab[:, :, 64:176, 48:144] = 1
    lab[1, :, 200:240, 10:250] = 1
    sdf = JF.signed_distance(D(lab.reshape(2, 256, 256), dev))
    for w in (5, 15):
        key = "%s_s%d_w%d" % (loss_type, loss_sum, w)
        x0 = big.clone().requires_grad_(True)
        ref = O.bev_head_loss(x0, lab, float(w), 20.0, 20.0, loss_type=loss_type, loss_sum=loss_sum)
        (g0,) = torch.autograd.grad(ref, x0)
        x1 = D(big, dev).requires_grad_(True)
        got = JF.bev_head_loss(x1, D(lab, dev), sdf, float(w), 20.0, 20.0, loss_type=loss_type, loss_sum=loss_sum)
        assert abs(got.item() - float(gold[key])) <= 3e-6 * max(abs(float(gold[key])), 1.0), key
        (g1,) = torch.autograd.grad(got, x1)
        assert (g1.cpu() - g0).abs().max().item() <= 2e-5 * g0.abs().max().item() + 1e-10, key
        assert abs(g1.abs().sum().item() - float(gold[key + "_gsum"])) <= 1e-4 * float(gold[key + "_gsum"]), key
