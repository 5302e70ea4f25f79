"""Host side of the tcgen05 implicit-GEMM convolution (``csrc/conv_tc.cu``): the per-layer K-chunk table,
weight packing for ragged channel counts, and the autograd binding.

Backward status (round 1): the forward runs on the hand-written kernel; the backward of each convolution still
calls the library (cuDNN through autograd on a re-materialised input).  dgrad / wgrad kernels are the next step
(DESIGN.md §roadmap)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr, stream_of

CL = torch.channels_last
ACT = {"none": 0, "relu": 1, "leaky": 2, "sigmoid": 3}
_TABLES: dict = {}


def _pad4(c):
    return (c + 3) // 4 * 4


def chunk_table(src_channels, kh, kw, device):
    """int32 [nkb*8, 4] table: one row per 16-byte chunk of K in weight order (tap-major, then source, then channel)."""
    key = (tuple(src_channels), kh, kw, str(device))
    t = _TABLES.get(key)
    if t is None:
        rows = []
        for ky in range(kh):
            for kx in range(kw):
                for si, c in enumerate(src_channels):
                    for coff in range(0, c, 4):
                        rows.append((si, (ky << 16) | (kx & 0xFFFF), coff, min(16, (c - coff) * 4)))
        while len(rows) % 8:
            rows.append((-1, 0, 0, 0))
        t = _TABLES[key] = torch.tensor(rows, dtype=torch.int32, device=device).contiguous()
    return t


def gemm_weight(weight, src_channels, weight_channels):
    """K-major GEMM operand [N, K] for the chunk order above.  The channels-last parameter is used in place when
    every source has a multiple of 4 channels and tensor channels == weight channels; otherwise a packed, zero-padded
    copy is built (Cin = 513 of iconv1-3, the 3/6-channel stems)."""
    N, Cin, kh, kw = weight.shape
    w = weight.permute(0, 2, 3, 1)   # [N, kh, kw, Cin] — the physical layout of a channels-last parameter
    if all(c % 4 == 0 for c in src_channels) and list(src_channels) == list(weight_channels):
        if not w.is_contiguous():
            w = w.contiguous()
        return w.reshape(N, kh * kw * Cin), kh * kw * Cin
    parts, off = [], 0
    for c_t, c_w in zip(src_channels, weight_channels):
        parts.append(F.pad(w[..., off:off + c_w], (0, _pad4(c_t) - c_w)))
        off += c_w
    wp = torch.cat(parts, -1).contiguous()
    return wp.reshape(N, -1), wp.shape[1] * wp.shape[2] * wp.shape[3]


def _torch_conv(xs, ups, weight, bias, stride, pad, reflect, act, residual):
    """Library formulation of the same operator (used for the interim backward and under host emulation)."""
    ts = [F.interpolate(t, scale_factor=2, mode="nearest") if up else t for t, up in zip(xs, ups)]
    x = ts[0] if len(ts) == 1 else torch.cat(ts, 1)
    if x.shape[1] > weight.shape[1]:     # zero-padded stem channels
        x = x[:, :weight.shape[1]]
    if reflect and pad:
        x = F.pad(x, (pad,) * 4, mode="reflect")
        pad = 0
    y = F.conv2d(x.contiguous(memory_format=CL), weight, bias, stride=stride, padding=pad)
    if residual is not None:
        y = y + residual
    if act == "relu":
        y = F.relu(y)
    elif act == "leaky":
        y = F.leaky_relu(y, 0.01)
    elif act == "sigmoid":
        y = torch.sigmoid(y)
    return y


class _ConvTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, weight, bias, residual, *xs):
        ups, stride, pad, reflect, act = cfg["ups"], cfg["stride"], cfg["pad"], cfg["reflect"], cfg["act"]
        xs = [x if x.is_contiguous(memory_format=CL) else x.contiguous(memory_format=CL) for x in xs]
        B = xs[0].shape[0]
        Hin = xs[0].shape[2] * (2 if ups[0] else 1)
        Win = xs[0].shape[3] * (2 if ups[0] else 1)
        N, Cin, kh, kw = weight.shape
        src_C = [x.shape[1] for x in xs]
        if len(xs) == 1 and src_C[0] != Cin:
            w_C = [Cin]
        else:
            w_C = src_C
            assert sum(src_C) == Cin, (src_C, Cin)
        Ho = (Hin + 2 * pad - kh) // stride + 1
        Wo = (Win + 2 * pad - kw) // stride + 1
        dev = xs[0].device
        table = chunk_table(src_C, kh, kw, dev)
        wmat, wcols = gemm_weight(weight.detach(), src_C, w_C)
        out = torch.empty((B, N, Ho, Wo), dtype=torch.float32, device=dev, memory_format=CL)
        a = _lib.ConvArgs()
        for i, x in enumerate(xs):
            a.src[i] = ptr(x)
            a.src_C[i], a.src_H[i], a.src_W[i], a.src_up[i] = x.shape[1], x.shape[2], x.shape[3], int(ups[i])
        a.nsrc = len(xs)
        a.B, a.Hin, a.Win, a.Ho, a.Wo, a.N = B, Hin, Win, Ho, Wo, N
        a.stride, a.pad, a.reflect = stride, pad, int(reflect)
        a.weight, a.w_row, a.w_cols = ptr(wmat), wmat.stride(0), wcols
        a.table, a.nkb = ptr(table), table.shape[0] // 8
        a.bias = ptr(bias.detach()) if bias is not None else None
        if residual is not None:
            residual = residual if residual.is_contiguous(memory_format=CL) else residual.contiguous(memory_format=CL)
            a.residual = ptr(residual)
        a.act = ACT[act]
        a.out = ptr(out)
        from .functional import _launch
        check(_launch("conv_fwd", out, lambda: _lib.lib().jpb_conv2d_fwd(C.byref(a), stream_of(out))), "jpb_conv2d_fwd")
        ctx.cfg = cfg
        ctx.has = (bias is not None, residual is not None)
        ctx.save_for_backward(weight, bias, residual, *xs)
        return out

    @staticmethod
    def backward(ctx, gy):
        cfg = ctx.cfg
        weight, bias, residual, *xs = ctx.saved_tensors
        with torch.enable_grad():
            xs_ = [x.detach().requires_grad_(True) for x in xs]
            w_ = weight.detach().requires_grad_(True)
            b_ = bias.detach().requires_grad_(True) if bias is not None else None
            r_ = residual.detach().requires_grad_(True) if residual is not None else None
            y = _torch_conv(xs_, cfg["ups"], w_, b_, cfg["stride"], cfg["pad"], cfg["reflect"], cfg["act"], r_)
            wanted = [w_] + ([b_] if b_ is not None else []) + ([r_] if r_ is not None else []) + xs_
            grads = list(torch.autograd.grad(y, wanted, gy, allow_unused=True))
        gw = grads.pop(0)
        gb = grads.pop(0) if b_ is not None else None
        gr = grads.pop(0) if r_ is not None else None
        return (None, gw, gb, gr) + tuple(grads)


def conv2d_tc(xs, ups, weight, bias, stride, pad, reflect, act, residual):
    cfg = dict(ups=tuple(bool(u) for u in ups), stride=stride, pad=pad, reflect=bool(reflect), act=act)
    return _ConvTC.apply(cfg, weight, bias, residual, *xs)
