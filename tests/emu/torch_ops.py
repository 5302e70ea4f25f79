"""Library (ATen / cuDNN) formulations of the fused operators — TEST INFRASTRUCTURE ONLY.

Two uses: (1) the reference side of the kernel parity tests (tests/test_conv.py, tests/test_heads.py, ...); (2) under the host
emulation of ``libjpb200.so`` the tcgen05/TMA convolution entry points do not exist, so ``install_conv_patch`` routes
``jperceiver_b200.conv.conv2d_tc`` through ``torch_conv`` WHILE THE EMULATION LIBRARY IS INSTALLED (and only then).  Nothing in
the ``jperceiver_b200`` package imports this module."""
from __future__ import annotations

import torch
import torch.nn.functional as F

CL = torch.channels_last


def torch_conv(xs, ups, weight, bias, stride, pad, reflect, act, residual):
    """nearest-2x up-sampling + channel concat + reflection/zero padding + conv2d + bias + residual + activation."""
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


def install_conv_patch():
    """Route ``conv.conv2d_tc`` through ``torch_conv`` while ``_lib.is_emulated()``; the real function otherwise."""
    from jperceiver_b200 import _lib, conv as JC
    if getattr(JC.conv2d_tc, "_emu_patch", False):
        return
    real = JC.conv2d_tc

    def conv2d_tc(xs, ups, weight, bias, stride, pad, reflect, act, residual, bn_stats=False):
        if _lib.is_emulated():
            JC.STATS_FUSED[0] = False
            return torch_conv(xs, ups, weight, bias, stride, pad, reflect, act, residual)
        return real(xs, ups, weight, bias, stride, pad, reflect, act, residual, bn_stats=bn_stats)

    conv2d_tc._emu_patch = True
    JC.conv2d_tc = conv2d_tc


def pose_head_torch(x, invert):
    """pose_decoder.py:22-26 + net.py:704-756: mean over the map, x0.01, Rodrigues, 4x4 assembly."""
    v = 0.01 * x.mean(3).mean(2)
    aa, t = v[:, :3], v[:, 3:]
    B = aa.shape[0]
    ang = aa.norm(dim=1, keepdim=True)
    ax = aa / (ang + 1e-7)
    ca, sa = torch.cos(ang)[:, 0], torch.sin(ang)[:, 0]
    Cc = 1 - ca
    x_, y_, z_ = ax[:, 0], ax[:, 1], ax[:, 2]
    R3 = torch.stack([x_ * x_ * Cc + ca, x_ * y_ * Cc - z_ * sa, z_ * x_ * Cc + y_ * sa,
                      x_ * y_ * Cc + z_ * sa, y_ * y_ * Cc + ca, y_ * z_ * Cc - x_ * sa,
                      z_ * x_ * Cc - y_ * sa, y_ * z_ * Cc + x_ * sa, z_ * z_ * Cc + ca], 1).view(B, 3, 3)
    R = torch.zeros(B, 4, 4, dtype=x.dtype, device=x.device)
    R[:, :3, :3] = R3
    R[:, 3, 3] = 1
    T = torch.eye(4, dtype=x.dtype, device=x.device).repeat(B, 1, 1)
    if invert:
        T[:, :3, 3] = -t
        return R.transpose(1, 2) @ T
    T[:, :3, 3] = t
    return T @ R


def cct_attention_torch(front, cross, front_hat, dfeat, p, conv2d):
    """CrossViewTransformer.py:45-92 after the depth-feature convs, with bmm / max / gather / broadcast matmul.
    ``conv2d``: the convolution to use around it (the product's, so only the attention core differs)."""
    B, C, a, b = front.shape
    n = a * b
    q = conv2d(cross, p.query_conv.weight, p.query_conv.bias).reshape(B, -1, n)
    k = conv2d(front, p.key_conv.weight, p.key_conv.bias).reshape(B, -1, n).permute(0, 2, 1)
    energy = torch.bmm(k, q)
    star, arg = energy.max(dim=1)
    v = conv2d(front_hat, p.value_conv.weight, p.value_conv.bias).reshape(B, -1, n)
    T = torch.gather(v, 2, arg.view(B, 1, n).expand(-1, v.shape[1], -1)).reshape(B, -1, a, b)
    S = star.view(B, 1, a, b)
    fused = conv2d([(front, False), (T, False)], p.f_conv.weight, p.f_conv.bias, pad=1)
    out = front + fused * S
    qd = conv2d(cross, p.query_conv_depth.weight, p.query_conv_depth.bias).reshape(B, -1, n)
    kd = conv2d(front, p.key_conv_depth.weight, p.key_conv_depth.bias).reshape(B, -1, n).permute(0, 2, 1)
    vd = conv2d(dfeat, p.value_conv_depth.weight, p.value_conv_depth.bias)
    attn = torch.bmm(kd, qd).max(dim=1)[0].view(B, 1, a, b)
    return (out + attn @ vd).contiguous(memory_format=CL), S, attn
