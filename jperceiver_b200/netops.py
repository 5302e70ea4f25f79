"""Network operators used by the model mirror (``jperceiver_b200.model``).

Activations are logical ``(B, C, H, W)`` tensors stored channels-last (physically NHWC), fp32.
Each operator below is one fused unit of the B200 design and runs as a hand-written sm_100a kernel
from ``libjpb200.so`` — there is one path: no library (cuDNN/ATen) formulation, no CPU path; every operator refuses
non-CUDA tensors and unsupported shapes raise.  (The library formulations the kernels are tested against live in
``tests/emu/torch_ops.py``.)
"""
from __future__ import annotations

import torch

from . import _lib
from . import conv as _conv

CL = torch.channels_last

# operator -> kernels (documentation; printed by bench.py)
KERNELS = {
    "conv2d": "csrc/conv_tc.cu (tcgen05 implicit GEMM: forward, dgrad, wgrad) + csrc/elementwise.cu + csrc/conv_smalln.cu",
    "maxpool": "csrc/pool.cu",
    "batchnorm": "csrc/bn.cu",
    "dropout": "csrc/heads.cu",
    "image_prep": "csrc/heads.cu",
    "pose_head": "csrc/heads.cu",
    "cvp_mlp": "csrc/heads.cu",
    "cct_attention": "csrc/heads.cu",
}


def _need_cuda(x):
    if not x.is_cuda and not _lib.is_emulated():
        raise _lib.JpbError("jperceiver_b200 runs on CUDA tensors only (got %s); there is no CPU fallback" % x.device)


FUSE_BN_STATS = True   # accumulate BatchNorm statistics in the epilogue of the convolution that feeds a training-mode BatchNorm


def conv2d(inputs, weight, bias=None, *, stride=1, pad=0, reflect=False, act="none", residual=None, bn_next=False):
    """Implicit-GEMM convolution with fused gather/epilogue.

    ``inputs``: a tensor, or a list of ``(tensor, up2x)`` pairs that are (nearest-2x up-sampled and)
    concatenated along channels on the fly — the ``cat(reduce(skip), upsample(x), disp)`` of
    depth_decoder.py:76-80 never exists in memory.  ``reflect``: ReflectionPad2d(pad) instead of zeros.
    Epilogue: ``+bias`` -> ``+residual`` -> activation.
    """
    if not isinstance(inputs, (list, tuple)):
        inputs = [(inputs, False)]
    _need_cuda(inputs[0][0])
    xs, ups = [t for t, _ in inputs], [u for _, u in inputs]
    y = _conv.conv2d_tc(xs, ups, weight, bias, stride, pad, reflect, act, residual, bn_stats=bool(bn_next and FUSE_BN_STATS))
    if bn_next and _conv.STATS_FUSED[0]:
        y._jpb_bn_stats = True      # consumed (and cleared) by the batchnorm() call that follows
    return y


def batchnorm(x, bn, training, *, relu=False, residual=None, momentum=0.1, eps=1e-5):
    """BatchNorm2d (+residual) (+ReLU).  Training mode uses per-GPU batch statistics and updates the
    running statistics in place, as nn.BatchNorm2d does."""
    _need_cuda(x)
    if x.shape[1] % 4:
        raise _lib.JpbError("batchnorm kernel needs a multiple of 4 channels (got %d)" % x.shape[1])
    from . import functional as JF
    if training:
        # ``stat_updates`` = 2 on the road-head BNs reproduces the reference's duplicated forward pass
        # (net.py:73-74): two momentum updates with the same batch statistics == one with 1-(1-m)^2.
        k = getattr(bn, "stat_updates", 1)
        nbt = bn.num_batches_tracked
        if not (torch.is_tensor(nbt) and nbt.is_cuda and nbt.dtype == torch.int64):
            bn.num_batches_tracked += k
            nbt = None
        ready = bool(getattr(x, "_jpb_bn_stats", False))
        if ready:
            x._jpb_bn_stats = False
        return JF.batchnorm_train(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                  1.0 - (1.0 - momentum) ** k, eps, relu, nbt, k, stats_ready=ready)
    return JF.batchnorm_eval(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, eps, relu)


def sum_n(xs):
    """Left-to-right sum of several equally shaped tensors in one pass (csrc/elementwise.cu: sum_n_kernel)."""
    _need_cuda(xs[0])
    from . import functional as JF
    if xs[0].numel() % 4:
        raise _lib.JpbError("sum_n kernel needs a multiple of 4 elements")
    return JF.sum_n(list(xs))


def maxpool(x, k, stride, pad):
    _need_cuda(x)
    if x.shape[1] % 4:
        raise _lib.JpbError("maxpool kernel needs a multiple of 4 channels (got %d)" % x.shape[1])
    from . import functional as JF
    return JF.maxpool(x, k, stride, pad)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, mask, seed, stream_id, step):
        from ._lib import check, ptr, stream_of
        xc = x.contiguous(memory_format=CL) if x.dim() == 4 else x.contiguous()
        y = torch.empty_like(xc)
        if mask is not None:
            mask = mask.to(torch.float32).expand_as(x)
            mask = mask.contiguous(memory_format=CL) if x.dim() == 4 else mask.contiguous()
        ctx.args = (float(p), mask, int(seed), int(stream_id), step)
        check(_lib.lib().jpb_dropout(ptr(xc), ptr(mask) if mask is not None else None, ptr(y), xc.numel(), float(p), int(seed),
                                     int(stream_id), ptr(step) if step is not None else None, stream_of(xc)), "jpb_dropout")
        return y

    @staticmethod
    def backward(ctx, gy):
        from ._lib import check, ptr, stream_of
        p, mask, seed, stream_id, step = ctx.args
        g = gy.contiguous(memory_format=CL) if gy.dim() == 4 else gy.contiguous()
        gx = torch.empty_like(g)
        check(_lib.lib().jpb_dropout(ptr(g), ptr(mask) if mask is not None else None, ptr(gx), g.numel(), p, seed, stream_id,
                                     ptr(step) if step is not None else None, stream_of(g)), "jpb_dropout(bwd)")
        return gx, None, None, None, None, None


_DROPOUT_CALLS = [0]


def dropout(x, p, training, mask=None, step=None, seed=0):
    """nn.Dropout; ``mask`` (0/1 keep tensor) overrides the random draw (parity tests).  The keep mask is a counter-based
    draw keyed on (seed, call index, device step counter, element): no mask tensor is stored, the backward regenerates it,
    and a replayed CUDA graph still sees a fresh mask every step."""
    if not training or p == 0.0:
        return x
    _need_cuda(x)
    _DROPOUT_CALLS[0] = (_DROPOUT_CALLS[0] + 1) % 4096
    return _Dropout.apply(x, p, mask, seed, 1000 + _DROPOUT_CALLS[0], step)


def image_prep(images, out_hw=None):
    """``(x - 0.45) / 0.225`` of one or two NCHW frames (concatenated along channels), optionally after a
    bilinear resize (align_corners=False), emitted channels-last with the channels zero-padded to 4 / 8."""
    if not isinstance(images, (list, tuple)):
        images = [images]
    _need_cuda(images[0])
    if len(images) == 1 and images[0].shape[1] == 6:     # an already concatenated frame pair (scripts/draw_odometry.py:66)
        images = [images[0][:, :3], images[0][:, 3:]]
    if len(images) > 2 or any(im.shape[1] != 3 or im.dtype != torch.float32 for im in images):
        raise _lib.JpbError("image_prep takes one or two 3-channel fp32 frames")
    from ._lib import check, ptr, stream_of
    ims = [im.contiguous() for im in images]
    B, _, Hs, Ws = ims[0].shape
    Ho, Wo = (Hs, Ws) if out_hw is None else tuple(out_hw)
    Cpad = 4 * len(ims)
    out = torch.empty((B, Cpad, Ho, Wo), dtype=torch.float32, device=ims[0].device, memory_format=CL)
    check(_lib.lib().jpb_image_prep(ptr(ims[0]), ptr(ims[1]) if len(ims) == 2 else None, ptr(out), B, Hs, Ws, Ho, Wo, Cpad,
                                    stream_of(out)), "jpb_image_prep")
    return out


def resize_bilinear(x, out_hw):
    """Bilinear resize (align_corners=False) of the CCT depth feature under the non-square rule: identity at
    the reference's 1024x1024; otherwise one ATen ``upsample_bilinear2d`` on a B x 512 x 10 x 32 map (the only library
    arithmetic left on the path)."""
    if tuple(x.shape[2:]) == tuple(out_hw):
        return x
    return torch.nn.functional.interpolate(x, list(out_hw), mode="bilinear", align_corners=False)


def _nhwc(t):
    return t if t.is_contiguous(memory_format=CL) else t.contiguous(memory_format=CL)


class _CvpMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2):
        from ._lib import check, ptr, stream_of
        x = _nhwc(x)
        B, Cc, h, w = x.shape
        n = h * w
        W1c, W2c = W1.detach().contiguous(), W2.detach().contiguous()
        y1, y2 = torch.empty_like(x, memory_format=CL), torch.empty_like(x, memory_format=CL)
        check(_lib.lib().jpb_cvp_mlp_fwd(ptr(x), ptr(W1c), ptr(b1.detach()), ptr(W2c), ptr(b2.detach()), ptr(y1), ptr(y2), B, n, Cc,
                                         stream_of(x)), "jpb_cvp_mlp_fwd")
        ctx.save_for_backward(x, W1c, W2c, y1, y2, W1, b1, W2, b2)
        return y2

    @staticmethod
    def backward(ctx, g):
        from ._lib import check, ptr, stream_of
        from .functional import direct_grad_target
        x, W1c, W2c, y1, y2, W1, b1, W2, b2 = ctx.saved_tensors
        B, Cc, h, w = x.shape
        g = _nhwc(g)
        dz2, dz1, dx = (torch.empty_like(x, memory_format=CL) for _ in range(3))
        targets = [direct_grad_target(p) for p in (W1, b1, W2, b2)]
        direct = all(t is not None and t.is_contiguous() for t in targets)
        dW1, db1, dW2, db2 = targets if direct else [torch.zeros_like(p) for p in (W1, b1, W2, b2)]
        check(_lib.lib().jpb_cvp_mlp_bwd(ptr(x), ptr(W1c), ptr(W2c), ptr(y1), ptr(y2), ptr(g), ptr(dz2), ptr(dz1), ptr(dx), ptr(dW1), ptr(db1),
                                         ptr(dW2), ptr(db2), B, h * w, Cc, stream_of(x)), "jpb_cvp_mlp_bwd")
        if direct:
            return dx, None, None, None, None
        return dx, dW1, db1, dW2, db2


def cvp_mlp(x, fc0, fc2):
    """Per-channel MLP over the flattened (h*w) positions: Linear+ReLU twice (CycledViewProjection.py:27-67)."""
    _need_cuda(x)
    if x.dtype != torch.float32:
        raise _lib.JpbError("cvp_mlp takes fp32 activations")
    return _CvpMlp.apply(x, fc0.weight, fc0.bias, fc2.weight, fc2.bias)


class _CctSelect(torch.autograd.Function):
    """(q, k, v, qd, kd) -> (T, S, attn): energies, hard max / arg-max over front positions, gather of the projected values."""

    @staticmethod
    def forward(ctx, q, k, v, qd, kd):
        from ._lib import check, ptr, stream_of
        q, k, v, qd, kd = (_nhwc(t) for t in (q, k, v, qd, kd))
        B, Cq, h, w = q.shape
        Cc, n = v.shape[1], h * w
        dev = q.device
        T = torch.empty_like(v, memory_format=CL)
        S = torch.empty(B, 1, h, w, dtype=torch.float32, device=dev)
        attn = torch.empty(B, 1, h, w, dtype=torch.float32, device=dev)
        arg = torch.empty(B, n, dtype=torch.int32, device=dev)
        argd = torch.empty(B, n, dtype=torch.int32, device=dev)
        check(_lib.lib().jpb_cct_select_fwd(ptr(q), ptr(k), ptr(v), ptr(qd), ptr(kd), ptr(T), ptr(S), ptr(arg), ptr(attn), ptr(argd), B, n, Cq, Cc,
                                            stream_of(q)), "jpb_cct_select_fwd")
        ctx.save_for_backward(q, k, qd, kd, arg, argd)
        ctx.geom = (B, n, Cq, Cc, h, w)
        ctx.mark_non_differentiable(arg, argd)
        return T, S, attn

    @staticmethod
    def backward(ctx, gT, gS, gattn):
        from ._lib import check, ptr, stream_of
        q, k, qd, kd, arg, argd = ctx.saved_tensors
        B, n, Cq, Cc, h, w = ctx.geom
        dev = q.device
        gT = _nhwc(gT) if gT is not None else torch.zeros((B, Cc, h, w), dtype=torch.float32, device=dev, memory_format=CL)
        gS = gS.contiguous() if gS is not None else torch.zeros(B, n, dtype=torch.float32, device=dev)
        gattn = gattn.contiguous() if gattn is not None else torch.zeros(B, n, dtype=torch.float32, device=dev)
        gq, gk, gqd, gkd = (torch.empty_like(q, memory_format=CL) for _ in range(4))
        gv = torch.empty((B, Cc, h, w), dtype=torch.float32, device=dev, memory_format=CL)
        check(_lib.lib().jpb_cct_select_bwd(ptr(q), ptr(k), ptr(qd), ptr(kd), ptr(arg), ptr(argd), ptr(gT), ptr(gS), ptr(gattn), ptr(gq), ptr(gk),
                                            ptr(gv), ptr(gqd), ptr(gkd), B, n, Cq, Cc, stream_of(q)), "jpb_cct_select_bwd")
        return gq, gk, gv, gqd, gkd


class _CctCombine(torch.autograd.Function):
    """out = front + fused * S + attn @ value_d."""

    @staticmethod
    def forward(ctx, front, fused, S, attn, vd):
        from ._lib import check, ptr, stream_of
        front, fused, vd = _nhwc(front), _nhwc(fused), _nhwc(vd)
        S, attn = S.contiguous(), attn.contiguous()
        B, Cc, h, w = front.shape
        out = torch.empty_like(front, memory_format=CL)
        check(_lib.lib().jpb_cct_combine_fwd(ptr(front), ptr(fused), ptr(S), ptr(attn), ptr(vd), ptr(out), B, h, w, Cc, stream_of(front)),
              "jpb_cct_combine_fwd")
        ctx.save_for_backward(fused, S, attn, vd)
        return out

    @staticmethod
    def backward(ctx, g):
        from ._lib import check, ptr, stream_of
        fused, S, attn, vd = ctx.saved_tensors
        B, Cc, h, w = fused.shape
        g = _nhwc(g)
        gfused, gvd = torch.empty_like(fused, memory_format=CL), torch.empty_like(vd, memory_format=CL)
        gS, gattn = torch.empty_like(S), torch.empty_like(attn)
        check(_lib.lib().jpb_cct_combine_bwd(ptr(g), ptr(fused), ptr(S), ptr(attn), ptr(vd), ptr(gfused), ptr(gS), ptr(gattn), ptr(gvd), B, h, w, Cc,
                                             stream_of(g)), "jpb_cct_combine_bwd")
        return g, gfused, gS, gattn, gvd


def cct_attention(front, cross, front_hat, dfeat, p):
    """Cross-view transformer core (CrossViewTransformer.py:45-92) after the depth-feature convs.
    Returns (out, S, attn).  Quirks kept: hard max/arg-max over keys; ``attn @ value_d`` is a batched
    (h x w) matrix product broadcast over channels."""
    _need_cuda(front)
    B, C, a, b = front.shape
    if a != b or a * b > 256:
        raise _lib.JpbError("cct_attention kernels take square maps of at most 256 positions (got %dx%d)" % (a, b))
    q = conv2d(cross, p.query_conv.weight, p.query_conv.bias)
    k = conv2d(front, p.key_conv.weight, p.key_conv.bias)
    v = conv2d(front_hat, p.value_conv.weight, p.value_conv.bias)
    qd = conv2d(cross, p.query_conv_depth.weight, p.query_conv_depth.bias)
    kd = conv2d(front, p.key_conv_depth.weight, p.key_conv_depth.bias)
    vd = conv2d(dfeat, p.value_conv_depth.weight, p.value_conv_depth.bias)
    T, S, attn = _CctSelect.apply(q, k, v, qd, kd)
    fused = conv2d([(front, False), (T, False)], p.f_conv.weight, p.f_conv.bias, pad=1)
    return _CctCombine.apply(front, fused, S, attn, vd), S, attn


class _PoseHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, invert):
        from ._lib import check, ptr, stream_of
        xc = x.contiguous(memory_format=CL)
        B, Cc, h, w = xc.shape
        T = torch.empty(B, 4, 4, dtype=torch.float32, device=x.device)
        mean6 = torch.empty(B, 6, dtype=torch.float32, device=x.device)
        check(_lib.lib().jpb_pose_head_fwd(ptr(xc), ptr(T), ptr(mean6), B, h * w, Cc, int(invert), stream_of(xc)), "jpb_pose_head_fwd")
        ctx.save_for_backward(mean6)
        ctx.geom = (B, Cc, h, w, int(invert))
        return T

    @staticmethod
    def backward(ctx, gT):
        from ._lib import check, ptr, stream_of
        (mean6,) = ctx.saved_tensors
        B, Cc, h, w, invert = ctx.geom
        gT = gT.contiguous()
        gx = torch.empty((B, Cc, h, w), dtype=torch.float32, device=gT.device, memory_format=CL)
        check(_lib.lib().jpb_pose_head_bwd(ptr(gT), ptr(mean6), ptr(gx), B, h * w, Cc, invert, stream_of(gT)), "jpb_pose_head_bwd")
        return gx, None


def pose_head(x, invert):
    """Spatial mean of the 6-channel PoseDecoder output, x0.01, Rodrigues, 4x4 assembly
    (pose_decoder.py:22-26, net.py:704-756)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.shape[1] < 6:
        raise _lib.JpbError("pose_head takes an fp32 map with at least 6 channels")
    return _PoseHead.apply(x, bool(invert))
