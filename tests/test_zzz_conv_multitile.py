"""Exact-operand parity of the tcgen05 convolution at extents where every resident CTA walks over SEVERAL tiles — the path the
persistent kernel takes at the benchmark shapes (TMEM accumulator double-buffering, the operand ring running across tile
boundaries, per-tile offset tables) and the wide-tile kernel takes with two CTAs per SM over several waves.  tests/test_conv.py
covers every layer flavour but at extents of at most a tile or two per CTA.  Operands are TF32-representable, so the only
difference from the fp32 library convolution is the summation order (tolerance 5e-5 of max|y|).
(Sorted last on purpose: the cases are the largest of the suite.)"""
import pytest
import torch

from jperceiver_b200 import _lib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu.torch_ops import torch_conv  # noqa: E402

from jperceiver_b200 import conv as JC

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def tf32(t):
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32).view(t.shape)


CASES = [
    # name, B, sources [(C, H, W, up)], Cout, k, stride, pad, reflect, act, bias
    ("persistent N=64: 3x3 64->64, 1024 tiles", 4, [(64, 128, 256, 0)], 64, 3, 1, 1, 0, "none", 0),
    ("persistent N=16: 3x3 up(16)->16, 1024 tiles", 2, [(16, 128, 128, 1)], 16, 3, 1, 1, 0, "none", 1),
    ("persistent stem: 7x7 s2 (3->pad4)->64, 512 tiles", 2, [(4, 256, 512, 0)], 64, 7, 2, 3, 0, "none", 0),
    ("wide tiles, 2 CTAs/SM: refl 3x3 256->256 leaky, 320 tiles", 2, [(256, 80, 256, 0)], 256, 3, 1, 1, 1, "leaky", 1),
    ("N=128, 2 CTAs/SM: 3x3 128->128, 512 tiles", 4, [(128, 128, 128, 0)], 128, 3, 1, 1, 0, "none", 0),
    # CTA pairs (tcgen05 cta_group::2, 256-wide tiles with more than 148 tiles): odd tile count (one idle partner), a ragged N
    # tile whose second half is partly outside the weight, and the three-source gather of the depth decoder's iconv layers
    ("CTA pairs, odd tile count: refl 3x3 64->256, 157 tiles", 2, [(64, 100, 100, 0)], 256, 3, 1, 1, 1, "none", 0),
    ("CTA pairs, N=192: 3x3 64->192 relu, 157 tiles", 2, [(64, 100, 100, 0)], 192, 3, 1, 1, 0, "relu", 1),
    ("CTA pairs, concat: refl 3x3 cat(128, up(128), 1)->256 leaky, 160 tiles", 4, [(128, 40, 128, 0), (128, 20, 64, 1), (1, 40, 128, 0)], 256, 3, 1, 1, 1,
     "leaky", 1),
    ("CTA pairs, 1x1 256->256, 160 tiles", 4, [(256, 40, 128, 0)], 256, 1, 1, 0, 0, "none", 1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_forward_many_tiles_per_cta(case):
    name, B, srcs, cout, k, stride, pad, reflect, act, has_bias = case
    _lib._handle, _lib._emulated = None, False
    pair = name.startswith("CTA pairs")
    if pair:      # the opt-in cta_group::2 schedule (csrc/conv_tc.cu: conv_pair), switched on for these cases only
        _lib.check(_lib.lib().jpb_conv_set_pair(1), "jpb_conv_set_pair")
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(len(name))
    xs = [tf32(torch.randn(B, c, h, w, generator=g)).to(dev).contiguous(memory_format=CL) for c, h, w, up in srcs]
    ups = [bool(up) for *_, up in srcs]
    cin_t = sum(c for c, *_ in srcs)
    cin_w = 3 if k == 7 else cin_t
    if k == 7:
        xs[0][:, cin_w:] = 0
    weight = tf32(torch.randn(cout, cin_w, k, k, generator=g) / (cin_w * k * k) ** 0.5).to(dev).contiguous(memory_format=CL)
    bias = torch.randn(cout, generator=g).to(dev) if has_bias else None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref0 = torch_conv(xs, ups, weight, bias, stride, pad, reflect, "none", None)
        ref = torch_conv(xs, ups, weight, bias, stride, pad, reflect, act, None)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    try:
        for rep in range(2):                      # twice: the second launch meets warm caches and a different block schedule
            with JC.trunc_comp(1.0):      # TF32-representable operands: nothing is truncated
                got = JC.conv2d_tc(xs, ups, weight, bias, stride, pad, reflect, act, None)
            torch.cuda.synchronize()
            err = (got - ref).abs().max().item()
            scale = max(ref0.abs().max().item(), 1e-6)
            assert err <= 5e-5 * scale, (name, rep, err, scale)
    finally:
        if pair:
            _lib.check(_lib.lib().jpb_conv_set_pair(0), "jpb_conv_set_pair")
