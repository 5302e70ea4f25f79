"""Device-side input pipeline against Pillow / torchvision — the libraries the reference's
``MonoDataset.preprocess`` calls — and against digests written from the reference's own ``preprocess`` run
(``oracle/make_golden_pipeline.py`` -> ``tests/golden/kat_pipeline.json``).  Byte / fixed-point work: everything is compared
BIT-EXACTLY (uint8 images, and the float tensors ``ToTensor`` makes of them).

``[gpu]`` calls libjpb200.so on cuda:0; ``[emu]`` runs the same kernel sources compiled as host C++.
(Sorted after the training-path suites: written after this round's GPU budget was spent; first on-device run.)"""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build_emulation  # noqa: E402
from oracle import pipeline_port as PP  # noqa: E402
from oracle.make_golden_pipeline import CASES  # noqa: E402

from jperceiver_b200 import _lib  # noqa: E402
from jperceiver_b200.datasets import preprocess as P  # noqa: E402

Image = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def dev(request):
    _lib._handle, _lib._emulated = None, False
    P._TABLE_CACHE.clear()
    if request.param == "emu":
        _lib.use_library(build_emulation(), emulated=True)
        yield torch.device("cpu")
    else:
        assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
        _lib.lib()
        yield torch.device("cuda:0")
    _lib._handle, _lib._emulated = None, False
    P._TABLE_CACHE.clear()


@pytest.mark.parametrize("h,w,oh,ow", [(37, 122, 37, 60), (370, 1226, 375, 1242), (375, 1242, 320, 1024), (50, 60, 128, 200),
                                       (64, 64, 64, 32), (33, 17, 11, 17), (20, 31, 20, 31)])
def test_lanczos_resize_bit_exact_with_pillow(dev, h, w, oh, ow):
    """Up-, down-sampling, skipped passes (equal width / height), per-sample flip; batch of two with different flags."""
    a = np.stack([PP.synth_frame(1, h, w), PP.synth_frame(2, h, w)])
    got, gf = P.resize_lanczos(torch.from_numpy(a).to(dev), (oh, ow), flip=[False, True])
    for b, flip in enumerate((False, True)):
        pil = Image.fromarray(a[b])
        if flip:
            pil = pil.transpose(Image.FLIP_LEFT_RIGHT)
        want = np.array(pil.resize((ow, oh), Image.LANCZOS))
        assert np.array_equal(got[b].cpu().numpy(), want)
        assert torch.equal(gf[b].cpu(), torch.from_numpy(want).permute(2, 0, 1).float().div(255))


def test_lanczos_extremes_clip_like_pillow(dev):
    """Saturated checkerboards drive the negative Lanczos lobes past 0 and 255 (clip8 lookups)."""
    a = np.zeros((2, 40, 64, 3), np.uint8)
    a[0, ::2, ::3] = 255
    a[1, :, 32:] = 255
    got, _ = P.resize_lanczos(torch.from_numpy(a).to(dev), (55, 90), want_float=False)
    for b in range(2):
        assert np.array_equal(got[b].cpu().numpy(), np.array(Image.fromarray(a[b]).resize((90, 55), Image.LANCZOS)))


def _jitter_ref(pil, order, fac, hue):
    from torchvision.transforms import _functional_pil as FP
    for op in order:
        pil = (FP.adjust_brightness(pil, fac[0]) if op == 0 else FP.adjust_contrast(pil, fac[1]) if op == 1
               else FP.adjust_saturation(pil, fac[2]) if op == 2 else FP.adjust_hue(pil, hue))
    return np.array(pil)


@pytest.mark.parametrize("order,fac,hue", [([0, 1, 2, 3], (0.83, 0.9, 0.81), 0.07), ([3, 2, 1, 0], (1.17, 1.13, 1.19), -0.093),
                                           ([2, 3, 0, 1], (1.2, 0.8, 1.2), 0.1), ([1, 0, 3, 2], (0.8, 1.2, 0.8), -0.1),
                                           ([3, 0, 1, 2], (1.0, 1.0, 1.0), 0.0)])
def test_color_jitter_bit_exact_with_torchvision_on_pil(dev, order, fac, hue):
    """Every operator order class (contrast first / last / in the middle: its mean is taken of the image as the preceding
    operators left it), factors on both sides of 1 (Blend.c's interpolate and extrapolate branches), hue wrap-around."""
    a = PP.synth_frame(5, 96, 160)
    a[:64, :64] = np.stack(list(np.meshgrid(np.arange(0, 256, 4), np.arange(0, 256, 4), indexing="ij")) + [np.full((64, 64), 77)], -1)
    got, gf = P.color_jitter(torch.from_numpy(a)[None].to(dev), torch.tensor([order]), torch.tensor([list(fac) + [0.0]]),
                             torch.tensor([int(hue * 255) & 255]))
    want = _jitter_ref(Image.fromarray(a), order, fac, hue)
    assert np.array_equal(got[0].cpu().numpy(), want)
    assert torch.equal(gf[0].cpu(), torch.from_numpy(want).permute(2, 0, 1).float().div(255))


def test_hue_round_trip_all_colours(dev):
    """Convert.c rgb2hsv / hsv2rgb over a 2^18-colour lattice (every 4th level per channel + the extremes) for three shifts;
    the full 2^24 cube was checked when the kernel was written (all five hue shifts, zero mismatches)."""
    from torchvision.transforms import _functional_pil as FP
    v = np.unique(np.concatenate([np.arange(0, 256, 4), [1, 2, 253, 254, 255]])).astype(np.uint8)
    a = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(len(v), -1, 3)
    for hue in (0.0, 0.07, -0.5):
        got = P.color_jitter(torch.from_numpy(a)[None].to(dev), torch.tensor([[3, 0, 1, 2]]), torch.tensor([[1.0, 1.0, 1.0, 0.0]]),
                             torch.tensor([int(hue * 255) & 255]), want_float=False)[0][0].cpu().numpy()
        assert np.array_equal(got, np.array(FP.adjust_hue(Image.fromarray(a), hue))), hue


def test_jitter_draws_follow_torchvision_rng_stream(dev):
    """``draw_color_jitter`` consumes torch's global RNG exactly as three calls of a ``transforms.ColorJitter`` module do."""
    import torchvision.transforms as T
    small = PP.synth_frame(9, 40, 56)
    torch.manual_seed(123)
    cj = T.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
    want = [np.array(cj(Image.fromarray(small))) for _ in range(3)]
    torch.manual_seed(123)
    order, factor, shift, _ = P.draw_color_jitter(3)
    got = P.color_jitter(torch.from_numpy(small)[None].repeat(3, 1, 1, 1).to(dev), order, factor, shift, enable=[1, 1, 0], want_float=False)[0]
    assert np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[1].cpu().numpy(), want[1])
    assert np.array_equal(got[2].cpu().numpy(), small)                        # do_color_aug off: pass-through


@pytest.mark.parametrize("h,w,size", [(1024, 1024, 256), (600, 777, 64), (100, 100, 256), (513, 511, 48), (257, 1023, 255)])
def test_bev_label_bit_exact_with_pillow(dev, h, w, size):
    a = np.stack([PP.synth_label(1, h, w), PP.synth_label(2, h, w)])
    a[0][np.random.RandomState(0).rand(h, w) > 0.97] = 128                     # grey levels are not road (== 255 test)
    got = P.bev_label(torch.from_numpy(a).to(dev), size, flip=[False, True]).cpu().numpy()
    assert np.array_equal(got[0, 0], PP.process_topview_both(Image.fromarray(a[0], "L"), size))
    assert np.array_equal(got[1, 0], PP.process_topview_both(Image.fromarray(a[1], "L"), size, flip=True))
    assert np.array_equal(got[1, 0], PP.process_topview(Image.fromarray(a[1], "L"), size, flip=True))   # two-level image


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_preprocess_matches_reference_digests(dev, case):
    """The whole colour path of ``MonoDataset.preprocess`` (flip -> resize_full -> resize -> ColorJitter -> ToTensor) for a
    three-frame snippet, against the digests of the reference's own run and the live oracle."""
    name, src, net, seed, do_aug, do_flip = case
    if name == "kitti_full" and dev.type == "cpu":
        pytest.skip("full-size case runs on the GPU (host emulation is one thread per block)")
    gold = json.load(open(os.path.join(GOLDEN, "kat_pipeline.json")))[name]
    frames = {f: torch.from_numpy(PP.synth_frame(10 * seed + j, *src))[None].to(dev) for j, f in enumerate([0, -1, 1])}
    pre = P.GpuPreprocess(net[0], net[1])
    torch.manual_seed(1000 + seed)
    params = [P.draw_color_jitter(3) if do_aug else None]
    out = pre(frames, [do_aug], [do_flip], params)
    import torchvision.transforms as T
    torch.manual_seed(1000 + seed)
    aug = T.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1)) if do_aug else None
    want = PP.preprocess_colour({f: Image.fromarray(PP.synth_frame(10 * seed + j, *src)) for j, f in enumerate([0, -1, 1])},
                                net[0], net[1], color_aug=aug, flip=do_flip)
    assert set(out) == set(want)
    for k, v in out.items():
        assert torch.equal(v[0].cpu(), want[k]), k
        assert PP.digest(v[0]) == gold["%s_%d_%d" % k], k
    lab = P.bev_label(torch.from_numpy(PP.synth_label(seed, 128, 128))[None].to(dev), net[0] // 4, flip=[do_flip])
    assert PP.digest(lab[0, 0]) == gold["bothS_0_0"]


def test_oracle_port_matches_reference_digests():
    """CPU only: the Pillow-based restatement against the digests of the reference's own ``preprocess``."""
    import torchvision.transforms as T
    gold = json.load(open(os.path.join(GOLDEN, "kat_pipeline.json")))
    for name, src, net, seed, do_aug, do_flip in CASES[:3]:
        torch.manual_seed(1000 + seed)
        aug = T.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1)) if do_aug else None
        got = PP.preprocess_colour({f: Image.fromarray(PP.synth_frame(10 * seed + j, *src)) for j, f in enumerate([0, -1, 1])},
                                   net[0], net[1], color_aug=aug, flip=do_flip)
        for k, v in got.items():
            assert PP.digest(v) == gold[name]["%s_%d_%d" % k], (name, k)
