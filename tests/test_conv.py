"""Layer-by-layer parity of the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) with the fp32 library
convolution on the shapes of SURVEY.md Appendix A (scaled down in batch/extent, same K/N structure).

Operands are drawn TF32-representable (low 13 mantissa bits zero), so the tensor-core products are exact and the only
difference from the fp32 library result is the fp32 summation order: forward tolerance 5e-5 of max|y| — a much sharper
check of indexing (gather table, swizzles, descriptors, scatter) than a TF32-noise tolerance would be, and activation
masks cannot flip between the two implementations.  A second forward test uses unrounded operands at 3e-3 (TF32)."""
import pytest
import torch
import torch.nn.functional as F

from jperceiver_b200 import _lib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu.torch_ops import torch_conv  # noqa: E402

from jperceiver_b200 import conv as JC

pytestmark = pytest.mark.gpu
JC.S2_MIN_PIXELS = 0          # the stride-2 parity-class data gradient at every test extent (the library only takes it for large layers)
CL = torch.channels_last


def tf32(t):
    """Truncate to TF32 (keep sign, 8 exponent and 10 mantissa bits)."""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32).view(t.shape)

CASES = [
    # name, sources [(C, H, W, up)], Cout, k, stride, pad, reflect, act, bias, residual
    ("layer1 3x3 64->64", [(64, 40, 64, 0)], 64, 3, 1, 1, 0, "none", 0, 0),
    ("layer2.0 3x3 s2 64->128", [(64, 40, 64, 0)], 128, 3, 2, 1, 0, "none", 0, 0),
    ("downsample 1x1 s2 64->128", [(64, 40, 64, 0)], 128, 1, 2, 0, 0, "none", 0, 0),
    ("layer3 3x3 256->256", [(256, 20, 32, 0)], 256, 3, 1, 1, 0, "none", 0, 0),
    ("layer4 3x3 512->512 (two N tiles)", [(512, 10, 16, 0)], 512, 3, 1, 1, 0, "none", 0, 0),
    ("stem 7x7 s2 (3->pad4)->64", [(4, 64, 96, 0)], 64, 7, 2, 3, 0, "none", 0, 0),
    ("pose stem 7x7 s2 (6->pad8)->64", [(8, 48, 80, 0)], 64, 7, 2, 3, 0, "none", 0, 0),
    ("reduce 1x1 128->256 no bias", [(128, 20, 64, 0)], 256, 1, 1, 0, 0, "none", 0, 0),
    ("crp 1x1 256->256 + residual", [(256, 20, 64, 0)], 256, 1, 1, 0, 0, "none", 0, 1),
    ("iconv4 refl 3x3 512->256 leaky", [(512, 10, 32, 0)], 256, 3, 1, 1, 1, "leaky", 1, 0),
    ("iconv3 refl 3x3 cat(256, up256, 1)->256 leaky", [(256, 20, 64, 0), (256, 10, 32, 1), (1, 20, 64, 0)], 256, 3, 1, 1, 1, "leaky", 1, 0),
    ("disp refl 3x3 up(256)->1 sigmoid", [(256, 20, 64, 1)], 1, 3, 1, 1, 1, "sigmoid", 1, 0),
    ("pose conv3 1x1 256->6", [(256, 6, 20, 0)], 6, 1, 1, 0, 0, "none", 1, 0),
    ("pose conv1 3x3 256->256 relu", [(256, 6, 20, 0)], 256, 3, 1, 1, 0, "relu", 1, 0),
    ("layout dec 3x3 up(32)->32", [(32, 32, 32, 1)], 32, 3, 1, 1, 0, "none", 1, 0),
    ("layout dec 3x3 16->16", [(16, 64, 64, 0)], 16, 3, 1, 1, 0, "none", 1, 0),
    ("topview refl 3x3 16->2", [(16, 64, 64, 0)], 2, 3, 1, 1, 1, "none", 1, 0),
    ("cvt f_conv 3x3 cat(128,128)->128", [(128, 8, 8, 0), (128, 8, 8, 0)], 128, 3, 1, 1, 0, "none", 1, 0),
    ("query 1x1 128->16", [(128, 8, 8, 0)], 16, 1, 1, 0, 0, "none", 1, 0),
    ("odd extent 3x3 64->64 (M not a multiple of 128)", [(64, 13, 19, 0)], 64, 3, 1, 1, 1, "none", 1, 0),
    # TMA-row A operand (JpbConvArgs.rows: rows of a multiple of 32 pixels): reflection at the row ends (31 + 1 pixel boxes), the data
    # gradient on the padded raster 258 -> 288, materialised up-sampled / padded sources, several zero-padded sources, a ragged tile
    ("rows: refl 3x3 128->128 leaky @24x256", [(128, 24, 256, 0)], 128, 3, 1, 1, 1, "leaky", 1, 0),
    ("rows: refl 3x3 cat(64, up64, 1)->128 leaky @16x256", [(64, 16, 256, 0), (64, 8, 128, 1), (1, 16, 256, 0)], 128, 3, 1, 1, 1, "leaky", 1, 0),
    ("rows: 1x1 64->128 + residual @15x96", [(64, 15, 96, 0)], 128, 1, 1, 0, 0, "none", 0, 1),
    ("rows: 3x3 cat(64, 64)->64 relu @16x64", [(64, 16, 64, 0), (64, 16, 64, 0)], 64, 3, 1, 1, 0, "relu", 1, 0),
    # data gradient of 3x3 / stride-2 layers as four parity-class launches (conv.py: _dgrad_s2_classes): TMA rows (W / 2 = 64) and
    # the gather kernels (W / 2 = 40, pose extent)
    ("s2 classes: layer3.0 3x3 s2 128->256 @40x128", [(128, 40, 128, 0)], 256, 3, 2, 1, 0, "none", 0, 0),
    ("s2 classes: pose layer2.0 3x3 s2 64->128 @48x80", [(64, 48, 80, 0)], 128, 3, 2, 1, 0, "none", 0, 0),
]


MODES = {
    # id: (operands rounded to TF32?, convolution precision, forward tolerance, backward tolerance) — tolerances relative to max|reference|
    "tf32-exact-operands": (True, "tf32", 5e-5, 3e-3),    # exact products: only the fp32 summation order differs (sharp indexing check)
    "fp32-operands": (False, "tf32", 3e-3, 3e-3),         # the benchmarked arithmetic: one TF32 product per term
    "3xtf32": (False, "3xtf32", 2e-5, 2e-5),              # split operands: fp32-grade results through the same kernels
}


@pytest.mark.parametrize("mode", list(MODES), ids=list(MODES))
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_forward_matches_fp32_library(case, mode):
    name, srcs, cout, k, stride, pad, reflect, act, has_bias, has_res = case
    exact, prec, tol, _ = MODES[mode]
    _lib._handle, _lib._emulated = None, False
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000)
    B = 2
    rnd = tf32 if exact else (lambda t: t)
    xs = [rnd(torch.randn(B, c, h, w, generator=g)).to(dev).contiguous(memory_format=CL) for c, h, w, up in srcs]
    ups = [bool(up) for *_, up in srcs]
    cin_t = sum(c for c, *_ in srcs)
    cin_w = {4: 3, 8: 6}.get(cin_t, cin_t) if k == 7 else cin_t
    if k == 7:
        xs[0][:, cin_w:] = 0
    weight = rnd(torch.randn(cout, cin_w, k, k, generator=g) / (cin_w * k * k) ** 0.5).to(dev).contiguous(memory_format=CL)
    bias = torch.randn(cout, generator=g).to(dev) if has_bias else None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref0 = torch_conv(xs, ups, weight, bias, stride, pad, reflect, "none", None)
        res = torch.randn(ref0.shape, generator=g).to(dev).contiguous(memory_format=CL) if has_res else None
        ref = torch_conv(xs, ups, weight, bias, stride, pad, reflect, act, res)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    # TF32-representable operands are not truncated by the tensor core: no truncation bias to compensate
    with JC.precision(prec), JC.trunc_comp(1.0 if exact else JC.TRUNC_COMP):
        got = JC.conv2d_tc(xs, ups, weight, bias, stride, pad, reflect, act, res)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and got.is_contiguous(memory_format=CL)
    err = (got - ref).abs().max().item()
    scale = max(ref0.abs().max().item(), 1e-6)
    assert err <= tol * scale, (name, mode, err, scale)


@pytest.mark.parametrize("mode", ["tf32-exact-operands", "3xtf32"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_backward_matches_fp32_library(case, mode):
    """dgrad (incl. reflection / up-sampling / concat scatter, stride-2 gather division), wgrad (MN-major tcgen05,
    split over pixels, packed K layouts) and the fused epilogue backward against fp32 autograd of the library form.
    ``3xtf32``: unrounded operands and upstream gradient, fp32-grade tolerance."""
    name, srcs, cout, k, stride, pad, reflect, act, has_bias, has_res = case
    exact, prec, _, tol = MODES[mode]
    rnd = tf32 if exact else (lambda t: t)
    _lib._handle, _lib._emulated = None, False
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000 + 7)
    B = 2
    xs = [rnd(torch.randn(B, c, h, w, generator=g)).to(dev).contiguous(memory_format=CL) for c, h, w, up in srcs]
    ups = [bool(up) for *_, up in srcs]
    cin_t = sum(c for c, *_ in srcs)
    cin_w = {4: 3, 8: 6}.get(cin_t, cin_t) if k == 7 else cin_t
    stem = k == 7
    if stem:
        xs[0][:, cin_w:] = 0
    weight = rnd(torch.randn(cout, cin_w, k, k, generator=g) / (cin_w * k * k) ** 0.5).to(dev).contiguous(memory_format=CL)
    bias = torch.randn(cout, generator=g).to(dev) if has_bias else None
    xm = [x.clone().requires_grad_(not stem) for x in xs]
    wm = weight.clone().requires_grad_(True)
    bm = bias.clone().requires_grad_(True) if has_bias else None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xr = [x.clone().requires_grad_(not stem) for x in xs]
        wr = weight.clone().requires_grad_(True)
        br = bias.clone().requires_grad_(True) if has_bias else None
        z = torch_conv(xr, ups, wr, br, stride, pad, reflect, "none", None)
        res = torch.randn(z.shape, generator=g).to(dev).contiguous(memory_format=CL) if has_res else None
        rr = res.clone().requires_grad_(True) if has_res else None
        rm = res.clone().requires_grad_(True) if has_res else None
        with JC.precision(prec), JC.trunc_comp(1.0 if exact else JC.TRUNC_COMP):
            got = JC.conv2d_tc(xm, ups, wm, bm, stride, pad, reflect, act, rm)
        ref = torch_conv(xr, ups, wr, br, stride, pad, reflect, act, rr)      # the reference's OWN activation masks
        gy = rnd(torch.randn(ref.shape, generator=g)).to(dev).contiguous(memory_format=CL)
        if act in ("relu", "leaky"):
            # the two summation orders differ by ~1e-5 of max|z|, so a handful of |z| ~ 0 elements may take different branches
            # of the activation; the upstream gradient is zeroed THERE (decided from the reference's pre-activation), which
            # makes the comparison independent of those branches without borrowing the kernel's masks
            zz = (z + rr).detach() if has_res else z.detach()
            gy = gy * (zz.abs() > 1e-4 * zz.abs().max()).to(gy.dtype)
        ref.backward(gy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    with JC.precision(prec), JC.trunc_comp(1.0 if exact else JC.TRUNC_COMP):
        got.backward(gy)
    torch.cuda.synchronize()

    def close(a, b, what):
        assert a is not None and a.shape == b.shape, (name, what)
        err = (a - b).abs().max().item()
        # tf32 mode: dz = gy * act'(y) is not TF32-representable after a leaky / sigmoid derivative: TF32 tolerance
        assert err <= tol * max(b.abs().max().item(), 1e-6), (name, mode, what, err, b.abs().max().item())

    close(wm.grad, wr.grad, "weight")
    if has_bias:
        close(bm.grad, br.grad, "bias")
    if has_res:
        close(rm.grad, rr.grad, "residual")
    if not stem:
        for i, (a, b) in enumerate(zip(xm, xr)):
            close(a.grad, b.grad, "input%d" % i)


@pytest.mark.gpu
@pytest.mark.parametrize("C,Cout,H,W,stride,bias", [(64, 64, 40, 64, 1, False), (64, 128, 40, 64, 2, False), (128, 16, 16, 16, 1, True),
                                                     (4, 64, 64, 96, 2, False)])
def test_bn_statistics_fused_into_conv_epilogue(C, Cout, H, W, stride, bias):
    """conv -> training BatchNorm with the statistics accumulated by the convolution epilogue (one BN launch) must equal the
    two-pass BatchNorm on the same convolution output: y, running statistics, num_batches_tracked, and all gradients."""
    from jperceiver_b200 import netops as ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, C, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    k = 7 if C == 4 else 3
    w = (torch.randn(Cout, C if C != 4 else 3, k, k, generator=g) * 0.1).to(dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, generator=g).to(dev) if bias else None
    res = {}
    slots, JC.KSPLIT_SLOTS = JC.KSPLIT_SLOTS, 0   # no split-K here: its atomic ordering makes two runs differ at TF32 level in the gradients
    for fuse in (False, True):
        ops.FUSE_BN_STATS = fuse
        bn = torch.nn.BatchNorm2d(Cout).to(dev).train()
        with torch.no_grad():
            bn.weight.copy_(torch.rand(Cout, generator=torch.Generator().manual_seed(1)) + 0.5)
            bn.bias.copy_(torch.randn(Cout, generator=torch.Generator().manual_seed(2)))
        xx, ww = x.clone().requires_grad_(C != 4), w.clone().requires_grad_(True)
        y = ops.conv2d(xx, ww, b, stride=stride, pad=k // 2, bn_next=True)
        fused = bool(getattr(y, "_jpb_bn_stats", False))
        assert not fused or fuse            # never fused when switched off
        Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
        nkb = (k * k * ((C + 3) // 4) + 7) // 8
        if fuse and JC._ksplit(4 * Ho * Wo, Cout, nkb) == 1:   # split-K launches cannot carry statistics (partial tiles)
            assert fused
        z = ops.batchnorm(y, bn, True, relu=True)
        gz = torch.randn(z.shape, generator=torch.Generator().manual_seed(3)).to(dev)
        grads = torch.autograd.grad(z, [ww, bn.weight, bn.bias] + ([xx] if C != 4 else []), gz)
        res[fuse] = [z.detach(), bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone().float()] + list(grads)
    ops.FUSE_BN_STATS = True
    JC.KSPLIT_SLOTS = slots
    for a_, b_ in zip(res[False], res[True]):
        assert (a_ - b_).abs().max().item() <= 1e-4 * max(a_.abs().max().item(), 1.0)   # fp32 partial sums in a different order
    # the shared accumulators are left clean: a plain two-pass BatchNorm right after gives the library result
    bn2 = torch.nn.BatchNorm2d(Cout).to(dev).train()
    t = torch.randn(2, Cout, 8, 8, device=dev).contiguous(memory_format=torch.channels_last)
    ref = torch.nn.functional.batch_norm(t, None, None, bn2.weight, bn2.bias, True)
    assert (ops.batchnorm(t, bn2, True) - ref).abs().max().item() < 1e-5


PATCH_CASES = [
    # C, N, H, W, act, bias, residual   (3x3 / stride 1 / zero pad 1 / one dense source: the TMA-patch kernel)
    (64, 64, 40, 64, "none", 0, 0),        # partial tile rows (40 = 2.5 x 16)
    (64, 64, 32, 32, "relu", 1, 1),
    (128, 128, 20, 64, "none", 0, 0),
    (256, 256, 20, 32, "leaky", 1, 0),
    (512, 512, 10, 16, "none", 0, 0),      # two N tiles, one tile row
    (32, 16, 64, 64, "none", 1, 0),
    (256, 64, 16, 8, "none", 0, 0),        # a single tile column
]


@pytest.mark.parametrize("tile_rows", [1, 2], ids=["one-tile", "two-tiles"])
@pytest.mark.parametrize("desc_mode", [0, 2], ids=["no-base-offset", "base-offset"])
@pytest.mark.parametrize("case", PATCH_CASES, ids=["%dx%d@%dx%d" % c[:4] for c in PATCH_CASES])
def test_conv_patch_kernel_forward_and_dgrad(case, desc_mode, tile_rows):
    """conv_tc_patch_kernel (A operand from 18x16 TMA patches, all nine taps on one patch through shifted UMMA descriptors)
    against the fp32 library convolution with TF32-representable operands: forward with the fused epilogue, and the data
    gradient (the same kernel on dz with flipped / transposed weights)."""
    import os
    want = int(os.environ.get("JPB_CONV_PATCH_DESC", "0"))
    if desc_mode != want and not os.environ.get("JPB_TEST_BOTH_DESC"):
        pytest.skip("descriptor variant %d (base offset set) is the measured-wrong experiment switch (JPB_TEST_BOTH_DESC=1 runs it)" % desc_mode)
    C, N, H, W, act, has_bias, has_res = case
    _lib._handle, _lib._emulated = None, False
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(C + N + H)
    B = 2
    x = tf32(torch.randn(B, C, H, W, generator=g)).to(dev).contiguous(memory_format=CL)
    weight = tf32(torch.randn(N, C, 3, 3, generator=g) / (C * 9) ** 0.5).to(dev).contiguous(memory_format=CL)
    bias = torch.randn(N, generator=g).to(dev) if has_bias else None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xr, wr = x.clone().requires_grad_(True), weight.clone().requires_grad_(True)
        z = torch_conv([xr], [False], wr, bias, 1, 1, False, "none", None)
        res = torch.randn(z.shape, generator=g).to(dev).contiguous(memory_format=CL) if has_res else None
        ref = torch_conv([xr], [False], wr, bias, 1, 1, False, act, res)
        gy = tf32(torch.randn(ref.shape, generator=g)).to(dev).contiguous(memory_format=CL)
        if act in ("relu", "leaky"):
            zz = (z + res).detach() if has_res else z.detach()
            gy = gy * (zz.abs() > 1e-4 * zz.abs().max()).to(gy.dtype)
        ref.backward(gy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    saved = JC.PATCH, JC.PATCH_MIN_TILES, JC.PATCH_DESC_MODE, JC.KSPLIT_SLOTS, JC.PATCH_TILE_ROWS
    JC.PATCH, JC.PATCH_MIN_TILES, JC.PATCH_DESC_MODE, JC.KSPLIT_SLOTS, JC.PATCH_TILE_ROWS = 1, 1, desc_mode, 0, tile_rows
    try:
        xm, wm = x.clone().requires_grad_(True), weight.clone().requires_grad_(True)
        with JC.trunc_comp(1.0):
            assert JC._patch_ok(C, N, H, W, B, 3, 3, 1, 1, False, 1, False, 1)
            got = JC.conv2d_tc([xm], [False], wm, bias, 1, 1, False, act, res)
            got.backward(gy)
        torch.cuda.synchronize()
    finally:
        JC.PATCH, JC.PATCH_MIN_TILES, JC.PATCH_DESC_MODE, JC.KSPLIT_SLOTS, JC.PATCH_TILE_ROWS = saved
    scale = max(z.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= 5e-5 * scale, ("forward", case, err, scale)
    errx = (xm.grad - xr.grad).abs().max().item()
    assert errx <= 3e-3 * max(xr.grad.abs().max().item(), 1e-6), ("dgrad", case, errx)
    errw = (wm.grad - wr.grad).abs().max().item()
    assert errw <= 3e-3 * max(wr.grad.abs().max().item(), 1e-6), ("wgrad", case, errw)


ROW_CASES = [
    # name, B, sources [(C, H, W, up)], Cout, k, reflect
    ("zero pad 3x3 64->64 @64x64", 2, [(64, 64, 64, 0)], 64, 3, 0),
    ("reflect 3x3 128->256 @24x128 (row-end boxes 31 + 1)", 2, [(128, 24, 128, 0)], 256, 3, 1),
    ("reflect 3x3 cat(64, up(64), 1)->128 @32x64 (gathered 1-channel source)", 2, [(64, 32, 64, 0), (64, 16, 32, 1), (1, 32, 64, 0)], 128, 3, 1),
    ("1x1 128->256 @32x32", 3, [(128, 32, 32, 0)], 256, 1, 0),
    ("zero pad 3x3 32->16 @16x96 (one group + dead groups)", 2, [(32, 16, 96, 0)], 16, 3, 0),
]


@pytest.mark.parametrize("case", ROW_CASES, ids=[c[0] for c in ROW_CASES])
def test_conv_wgrad_tma_rows_matches_gather_and_library(case):
    """Weight gradient with the im2col^T operand fetched by TMA boxes (csrc/conv_tc.cu: conv_tc_wgrad_kernel<.., ROWS>, the default
    for stride-1 same-size convolutions with rows of a multiple of 32 pixels) against the gathered operand and against fp32
    autograd of the library convolution; TF32-representable operands, so only the summation order differs."""
    name, B, srcs, cout, k, reflect = case
    _lib._handle, _lib._emulated = None, False
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(len(name))
    xs = [tf32(torch.randn(B, c, h, w, generator=g)).to(dev).contiguous(memory_format=CL) for c, h, w, up in srcs]
    ups = [bool(up) for *_, up in srcs]
    cin = sum(c for c, *_ in srcs)
    pad = (k - 1) // 2
    weight = tf32(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).contiguous(memory_format=CL)
    H, W = srcs[0][1], srcs[0][2]
    dz = tf32(torch.randn(B, cout, H, W, generator=g)).to(dev).contiguous(memory_format=CL)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        wr = weight.clone().requires_grad_(True)
        torch_conv(xs, ups, wr, None, 1, pad, reflect, "none", None).backward(dz)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    got = {}
    saved = JC.WGRAD_ROWS
    try:
        for rows in (0, 1):
            JC.WGRAD_ROWS = rows
            with JC.trunc_comp(1.0):
                got[rows] = JC.conv_wgrad(dz, weight, xs, ups, 1, pad, reflect).contiguous()
            torch.cuda.synchronize()
    finally:
        JC.WGRAD_ROWS = saved
    scale = wr.grad.abs().max().item()
    assert (got[0] - wr.grad).abs().max().item() <= 1e-4 * scale, (name, "gather vs library")
    assert (got[1] - wr.grad).abs().max().item() <= 1e-4 * scale, (name, "rows vs library")
