"""Autograd bindings of the C-ABI kernels (``include/jpb200.h``): one ``torch.autograd.Function`` per
fused operator.  PyTorch is used for device memory, streams and the autograd tape only; all arithmetic
happens in ``libjpb200.so``."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import PhotoArgs, PhotoGrad, check, ptr, stream_of


# ---- optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline figures) ----
PROFILE_ON = False
PROFILE: dict = {}
PROFILE_DETAIL: dict = {}     # (name, tag) -> events; tag = the caller's shape description (per-layer tables, tools/conv_layers.py)
# An eager step is host-bound (~40 us of Python / ctypes / tensor-map encoding per launch against 5-50 us kernels): with an idle
# GPU the first event of a pair is stamped when the host reaches it and the pair measures the host's launch latency, not the
# kernel.  PROFILE_BACKPRESSURE > 0 queues a device spin of that many SM cycles every PROFILE_EVERY launches so that the GPU
# always runs behind the host and each event pair brackets device time only.
PROFILE_BACKPRESSURE = 0
PROFILE_EVERY = 48
_profile_count = [0]


def _launch(name, tensor, call, tag=None):
    """Run one C-ABI launch; when profiling, bracket it with CUDA events on the tensor's current stream."""
    if PROFILE_ON and tensor.is_cuda:
        if PROFILE_BACKPRESSURE:
            if _profile_count[0] % PROFILE_EVERY == 0:
                torch.cuda._sleep(int(PROFILE_BACKPRESSURE))
            _profile_count[0] += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        status = call()
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1))
        if tag is not None:
            PROFILE_DETAIL.setdefault((name, tag), []).append((e0, e1))
        return status
    return call()


def profile_summary():
    """{kernel: {launches, ms_per_launch, ms_total}} — call after a device synchronize."""
    out = {}
    for name, evs in PROFILE.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        out[name] = {"launches": len(ms), "ms_per_launch": sum(ms) / len(ms), "ms_total": sum(ms)}
    return out


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def finalize(acc, scale, den=None):
    """float32 tensor = acc / den * scale, on device (no host sync)."""
    out = torch.empty(acc.shape, dtype=torch.float32, device=acc.device)
    check(_lib.lib().jpb_finalize(ptr(acc), ptr(den), float(scale), ptr(out), acc.numel(), stream_of(acc)), "jpb_finalize")
    return out


# --------------------------------------------------------------------------------------------------
# fused photometric reprojection loss (one launch per scale)
# --------------------------------------------------------------------------------------------------
def _photo_args(disp, target, sources, Ts, noises, K, invK, automask, min_depth, max_depth, noise_scale, seed, stream_id, step=None):
    a = PhotoArgs()
    B, _, H, W = target.shape
    a.target = ptr(target)
    for i, (s, T) in enumerate(zip(sources, Ts)):
        a.src[i] = ptr(s)
        a.T[i] = ptr(T)
        a.noise[i] = ptr(noises[i]) if noises is not None else None
    a.disp, a.K, a.invK = ptr(disp), ptr(K), ptr(invK)
    a.B, a.H, a.W, a.hs, a.ws, a.F = B, H, W, disp.shape[-2], disp.shape[-1], len(sources)
    a.automask = int(bool(automask))
    a.min_disp, a.max_disp = 1.0 / max_depth, 1.0 / min_depth
    a.noise_scale = float(noise_scale)
    a.seed, a.stream = int(seed), int(stream_id)
    a.step = ptr(step) if step is not None else None
    return a


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, K, invK, target, cfg, *rest):
        F = cfg["F"]
        sources, Ts = rest[:F], rest[F:2 * F]
        noises = rest[2 * F:3 * F] if len(rest) > 2 * F else None
        disp_c, K, invK, target = _f32c(disp), _f32c(K), _f32c(invK), _f32c(target)
        sources = [_f32c(s) for s in sources]
        Ts = [_f32c(T) for T in Ts]
        if noises is not None:
            noises = [_f32c(n) for n in noises]
        B, _, H, W = target.shape
        dev = target.device
        a = _photo_args(disp_c, target, sources, Ts, noises, K, invK, cfg["automask"], cfg["min_depth"],
                        cfg["max_depth"], cfg["noise_scale"], cfg["seed"], cfg["stream"], cfg.get("step"))
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        winner = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
        a.loss_sum, a.winner = ptr(acc), ptr(winner)
        min_index, warped = None, []
        if cfg["debug_outputs"]:
            min_index = torch.empty(B, H, W, dtype=torch.int64, device=dev)
            a.min_index = ptr(min_index)
        keep = bool(cfg.get("keep_warped", False)) and (ctx.needs_input_grad[0] or any(ctx.needs_input_grad[5 + F:5 + 2 * F]))
        if cfg["debug_outputs"] or keep:
            for i in range(F):
                w = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
                warped.append(w)
                a.warped[i] = ptr(w)
        cache = cfg.get("ident_cache")
        if cache is not None and cfg["automask"] and F <= 2 and _lib.photo_fwd_variant() == 3:
            # the identity candidates do not depend on the scale (net.py:159-166): the first launch of a step stores their
            # un-noised errors, the other scales' launches read them instead of evaluating two more SSIM candidates
            key = (target.data_ptr(), tuple(s.data_ptr() for s in sources), B, H, W)
            if cache.get("key") == key and cache.get("err") is not None:
                a.ident_err, a.ident_mode = ptr(cache["err"]), 2
            else:
                cache["err"] = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev)
                cache["key"] = key
                a.ident_err, a.ident_mode = ptr(cache["err"]), 1
        check(_launch("photometric_fwd", target, lambda: _lib.lib().jpb_photometric_fwd(C.byref(a), stream_of(target))),
              "jpb_photometric_fwd")
        loss = finalize(acc, 1.0 / (B * H * W * cfg["num_scales"])).reshape(())
        ctx.cfg = cfg
        ctx.noises = noises
        ctx.kept = len(warped) if keep else 0
        ctx.save_for_backward(disp_c, K, invK, target, winner, *sources, *Ts, *(warped if keep else []))
        ctx.mark_non_differentiable(winner)
        outs = [loss, winner]
        if cfg["debug_outputs"]:
            ctx.mark_non_differentiable(min_index, *warped)
            outs += [min_index] + warped
        return tuple(outs)

    @staticmethod
    def backward(ctx, gloss, *unused):
        cfg = ctx.cfg
        F = cfg["F"]
        disp, K, invK, target, winner = ctx.saved_tensors[:5]
        sources, Ts = ctx.saved_tensors[5:5 + F], ctx.saved_tensors[5 + F:5 + 2 * F]
        B, _, H, W = target.shape
        a = _photo_args(disp, target, sources, Ts, ctx.noises, K, invK, cfg["automask"], cfg["min_depth"],
                        cfg["max_depth"], cfg["noise_scale"], cfg["seed"], cfg["stream"])
        for i, w in enumerate(ctx.saved_tensors[5 + 2 * F:5 + 2 * F + ctx.kept]):
            a.warped[i] = ptr(w)          # phase A of the backward stages these instead of re-projecting the apron
        g = PhotoGrad()
        gl = _f32c(gloss).reshape(1)
        gdisp = torch.zeros_like(disp)
        gT = [torch.zeros_like(T) for T in Ts]
        g.grad_out, g.inv_count, g.winner, g.grad_disp = ptr(gl), 1.0 / (B * H * W * cfg["num_scales"]), ptr(winner), ptr(gdisp)
        for i in range(F):
            g.grad_T[i] = ptr(gT[i])
        check(_launch("photometric_bwd", target, lambda: _lib.lib().jpb_photometric_bwd(C.byref(a), C.byref(g), stream_of(target))),
              "jpb_photometric_bwd")
        n_extra = len(ctx.noises) if ctx.noises is not None else 0
        return (gdisp, None, None, None, None) + (None,) * F + tuple(gT) + (None,) * n_extra


# The backward stages the warped frames its forward wrote instead of re-projecting the 2-pixel apron: measured -9 % on the backward
# (profiles/r2_photometric_ab.jsonl); with the drop-in's default outputs (("color", f, s)) those frames exist anyway.
KEEP_WARPED = os.environ.get("JPB_PHOTO_KEEP_WARPED", "1") not in ("", "0")


def photometric_loss(disp, target, sources, Ts, K, invK, *, num_scales=4, automask=True, min_depth=0.1,
                     max_depth=100.0, noise=None, noise_scale=1e-5, seed=0, stream=0, step=None, debug_outputs=False,
                     keep_warped=None, ident_cache=None):
    """``loss_dict[("min_reconstruct_loss", s)]`` of one scale (already divided by ``num_scales``).

    Returns ``(loss, winner_u8, min_index|None, [warped...])``.  ``noise``: list of B×H×W tensors for the
    identity terms (tests) or None for the in-kernel Philox draw scaled by ``noise_scale``.  ``ident_cache``: a dict shared by
    the per-scale calls of ONE step over the same target / sources — the first call stores the (scale-independent) identity
    errors in it, the others read them (``JpbPhotoArgs.ident_mode``).
    """
    cfg = dict(F=len(sources), num_scales=num_scales, automask=automask, min_depth=min_depth, max_depth=max_depth,
               noise_scale=noise_scale, seed=seed, stream=stream, step=step, debug_outputs=debug_outputs,
               keep_warped=KEEP_WARPED if keep_warped is None else bool(keep_warped), ident_cache=ident_cache)
    rest = list(sources) + list(Ts) + (list(noise) if noise is not None else [])
    out = _Photometric.apply(disp, K, invK, target, cfg, *rest)
    loss, winner = out[0], out[1]
    if debug_outputs:
        return loss, winner, out[2], list(out[3:])
    return loss, winner, None, []


# --------------------------------------------------------------------------------------------------
# area pyramid + smoothness
# --------------------------------------------------------------------------------------------------
def area_pyramid(img, nlev):
    """[F.interpolate(img, /2^(s+1), mode='area') for s in range(nlev)] in one pass (no gradient)."""
    img = _f32c(img)
    B, Cc, H, W = img.shape
    out = _lib.Pyramid()
    out.nlev = nlev
    levels = []
    for s in range(nlev):
        t = torch.empty(B, Cc, H >> (s + 1), W >> (s + 1), dtype=torch.float32, device=img.device)
        levels.append(t)
        out.level[s] = ptr(t)
    check(_lib.lib().jpb_area_pyramid(ptr(img), B * Cc, H, W, C.byref(out), stream_of(img)), "jpb_area_pyramid")
    return levels


class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, J, weight, disp_norm):
        disp_c, J = _f32c(disp), _f32c(J)
        B, _, h, w = disp_c.shape
        acc = torch.zeros(B, 7, dtype=torch.float64, device=disp_c.device)
        out = torch.empty((), dtype=torch.float32, device=disp_c.device)
        check(_lib.lib().jpb_smooth_fwd(ptr(disp_c), ptr(J), B, h, w, int(disp_norm), float(weight), ptr(acc), ptr(out),
                                        stream_of(disp_c)), "jpb_smooth_fwd")
        ctx.save_for_backward(disp_c, J, acc)
        ctx.args = (float(weight), int(disp_norm))
        return out

    @staticmethod
    def backward(ctx, g):
        disp, J, acc = ctx.saved_tensors
        weight, disp_norm = ctx.args
        B, _, h, w = disp.shape
        gd = torch.zeros_like(disp)
        check(_lib.lib().jpb_smooth_bwd(ptr(disp), ptr(J), B, h, w, disp_norm, weight, ptr(acc), ptr(_f32c(g).reshape(1)), ptr(gd),
                                        stream_of(disp)), "jpb_smooth_bwd")
        return gd, None, None, None


def smooth_loss(disp, J, weight, disp_norm=True):
    """``loss_dict[("smooth_loss", s)]`` (weight = smoothness_weight / 2^s / num_scales already applied)."""
    return _Smooth.apply(disp, J, weight, disp_norm)


# --------------------------------------------------------------------------------------------------
# CGT scale label / loss
# --------------------------------------------------------------------------------------------------
def scale_label(label, odometry_K, Tr, out_hw, *, split, mode, quad=None, align_corners=True, occ=None):
    """CGT scale label.  ``mode``: 'both' (net.py:403-476), 'static' (:212-310) or 'dynamic' (:311-402; the reference
    reads the BEV label there only for its shape, so ``label`` may be None when ``occ`` is given)."""
    Kc, Tr = _f32c(odometry_K), _f32c(Tr)
    if mode != "dynamic" or label is not None:
        label = _f32c(label)
        occ = label.shape[-1]
        if label.shape[-2] != occ:
            raise ValueError("The shape of both label is not %d" % occ)   # net.py:222-223
    B = Kc.shape[0]
    Hf, Wf = out_hw
    out = torch.empty(B, 1, Hf, Wf, dtype=torch.float32, device=Kc.device)
    a = _lib.ScaleLabelArgs()
    a.label, a.K3, a.Tr, a.out = (ptr(label) if label is not None else None), ptr(Kc), ptr(Tr), ptr(out)
    a.k_stride, a.k_row = Kc.shape[-2] * Kc.shape[-1], Kc.shape[-1]
    a.quad = ptr(quad) if quad is not None else None
    a.B, a.occ, a.Hf, a.Wf = B, occ, Hf, Wf
    a.mode = {"both": 0, "static": 1, "dynamic": 2}[mode]
    a.align_corners = int(align_corners)
    a.z_offset = 1.9 if split == "argo" else (0.0 if mode == "dynamic" else 0.27)
    a.cam_height = 0.33 if split == "argo" else 1.73
    check(_lib.lib().jpb_scale_label(C.byref(a), stream_of(Kc)), "jpb_scale_label")
    return out


class _ScaleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, label, weight, crop, min_depth, max_depth):
        disp_c, label = _f32c(disp), _f32c(label)
        acc = torch.zeros(2, dtype=torch.float64, device=disp_c.device)
        a = _ScaleLoss._args(disp_c, label, acc, weight, crop, min_depth, max_depth)
        check(_lib.lib().jpb_scale_loss_fwd(C.byref(a), stream_of(disp_c)), "jpb_scale_loss_fwd")
        ctx.save_for_backward(disp_c, label, acc)
        ctx.args = (weight, crop, min_depth, max_depth)
        return finalize(acc[:1], weight, acc[1:]).reshape(())

    @staticmethod
    def _args(disp, label, acc, weight, crop, min_depth, max_depth):
        a = _lib.ScaleLossArgs()
        a.disp, a.label, a.acc = ptr(disp), ptr(label), ptr(acc)
        a.B, a.hs, a.ws, a.Hf, a.Wf = disp.shape[0], disp.shape[-2], disp.shape[-1], label.shape[-2], label.shape[-1]
        a.crop = int(crop)
        a.min_disp, a.max_disp = 1.0 / max_depth, 1.0 / min_depth
        a.weight = float(weight)
        return a

    @staticmethod
    def backward(ctx, g):
        disp, label, acc = ctx.saved_tensors
        a = _ScaleLoss._args(disp, label, acc, *ctx.args)
        gd = torch.zeros_like(disp)
        gl = _f32c(g).reshape(1)
        a.grad_out, a.grad_disp = ptr(gl), ptr(gd)
        check(_lib.lib().jpb_scale_loss_bwd(C.byref(a), stream_of(disp)), "jpb_scale_loss_bwd")
        return gd, None, None, None, None, None


def scale_loss(disp, label, weight, crop=False, min_depth=0.1, max_depth=100.0):
    """``loss_dict[("scale_loss", s)]`` (weight = scale_weight / 2^s / num_scales already applied)."""
    return _ScaleLoss.apply(disp, label, float(weight), bool(crop), min_depth, max_depth)


# --------------------------------------------------------------------------------------------------
# BEV head losses
# --------------------------------------------------------------------------------------------------
def signed_distance(label):
    """Exact signed distance map of every B×occ×occ binary label map, on device."""
    label = _f32c(label)
    B, n = label.shape[0], label.shape[-1]
    work = torch.empty(2 * B * n * n, dtype=torch.int32, device=label.device)
    sdf = torch.empty(B, n, n, dtype=torch.float32, device=label.device)
    check(_lib.lib().jpb_signed_distance(ptr(label), B, n, ptr(work), ptr(sdf), stream_of(label)), "jpb_signed_distance")
    return sdf


class _BevLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, sdf, w_fg, lw, l2w, region=0, loss_sum=3):
        if logits.dtype != torch.float32:
            logits = logits.float()
        if not (logits.is_contiguous() or logits.is_contiguous(memory_format=torch.channels_last)):
            logits = logits.contiguous()
        label = _f32c(label)
        B, _, occ, _ = logits.shape
        acc = torch.zeros(4 * B + 4, dtype=torch.float64, device=logits.device)
        out = torch.empty((), dtype=torch.float32, device=logits.device)
        a = _BevLoss._args(logits, label, sdf, acc, w_fg, lw, l2w, region, loss_sum)
        check(_lib.lib().jpb_bev_loss_fwd(C.byref(a), ptr(out), stream_of(logits)), "jpb_bev_loss_fwd")
        ctx.save_for_backward(logits, label, sdf, acc)
        ctx.args = (w_fg, lw, l2w, region, loss_sum)
        return out

    @staticmethod
    def _args(logits, label, sdf, acc, w_fg, lw, l2w, region=0, loss_sum=3):
        a = _lib.BevArgs()
        sb, sc, sy, sx = logits.stride()
        assert sy == logits.shape[3] * sx, "logits rows must be dense"
        a.logits, a.stride_b, a.stride_c, a.stride_p = ptr(logits), sb, sc, sx
        a.label, a.sdf, a.acc = ptr(label), ptr(sdf), ptr(acc)
        a.B, a.occ = logits.shape[0], logits.shape[2]
        a.w_fg, a.loss_weight, a.loss2_weight = float(w_fg), float(lw), float(l2w)
        a.region, a.loss_sum = int(region), int(loss_sum)
        return a

    @staticmethod
    def backward(ctx, g):
        logits, label, sdf, acc = ctx.saved_tensors
        a = _BevLoss._args(logits, label, sdf, acc, *ctx.args)
        gl = torch.empty_like(logits)  # preserves the (dense) strides
        check(_lib.lib().jpb_bev_loss_bwd(C.byref(a), ptr(_f32c(g).reshape(1)), ptr(gl), stream_of(logits)), "jpb_bev_loss_bwd")
        return gl, None, None, None, None, None, None, None


BEV_REGION = {"iou": 0, "dice": 1, "tversky": 2, "focal": 3}


def bev_head_loss(logits, label, sdf, w_fg, loss_weight=20.0, loss2_weight=20.0, loss_type="iou", loss_sum=3):
    """``compute_topview_loss`` (net.py:554-617): ``loss_type`` in {iou, dice, tversky, focal}; ``loss_sum`` 1 = region term
    only, 2 = + boundary loss, 3 = + boundary loss + weighted cross entropy (``loss2_type`` is 'boundary' in every config)."""
    if loss_type not in BEV_REGION:
        raise NotImplementedError("loss_type %r (the reference defines iou, dice, focal, tversky)" % (loss_type,))
    if loss_sum not in (1, 2, 3):
        raise NotImplementedError("loss_sum %r (the reference defines 1, 2, 3)" % (loss_sum,))
    return _BevLoss.apply(logits, label, sdf, w_fg, loss_weight, loss2_weight, BEV_REGION[loss_type], loss_sum)


class _L1Mean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _f32c(x), _f32c(y)
        acc = torch.zeros(1, dtype=torch.float64, device=x.device)
        check(_lib.lib().jpb_l1_mean_fwd(ptr(x), ptr(y), x.numel(), ptr(acc), stream_of(x)), "jpb_l1_mean_fwd")
        ctx.save_for_backward(x, y)
        return finalize(acc, 1.0 / x.numel()).reshape(())

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        check(_lib.lib().jpb_l1_mean_bwd(ptr(x), ptr(y), x.numel(), ptr(_f32c(g).reshape(1)), ptr(gx), ptr(gy), stream_of(x)),
              "jpb_l1_mean_bwd")
        return gx, gy


def l1_mean(x, y):
    """``nn.L1Loss()(x, y)`` (compute_transform_losses).  x and y must share a memory layout."""
    return _L1Mean.apply(x, y)


# --------------------------------------------------------------------------------------------------
# left-to-right sum of several tensors in one pass (the residual chain of a CRP block)
# --------------------------------------------------------------------------------------------------
class _SumN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        cl = torch.channels_last
        xs = [x if x.is_contiguous(memory_format=cl) else x.contiguous(memory_format=cl) for x in xs]
        y = torch.empty_like(xs[0], memory_format=cl)
        arr = (C.c_void_p * len(xs))(*[ptr(x) for x in xs])
        check(_launch("sum_n", y, lambda: _lib.lib().jpb_sum_n(arr, len(xs), ptr(y), y.numel(), stream_of(y))), "jpb_sum_n")
        ctx.n = len(xs)
        return y

    @staticmethod
    def backward(ctx, g):
        return (g,) * ctx.n


def sum_n(xs):
    """``((xs[0] + xs[1]) + xs[2]) + ...`` of 2..8 equally shaped channels-last tensors, bit-identical to the chain of binary adds."""
    return _SumN.apply(*xs)


# --------------------------------------------------------------------------------------------------
# NHWC max pooling
# --------------------------------------------------------------------------------------------------
class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p):
        x = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)
        B, Cc, H, W = x.shape
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        y = torch.empty((B, Cc, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty((B, Ho, Wo, Cc), dtype=torch.uint8, device=x.device)
        check(_launch("maxpool_fwd", x, lambda: _lib.lib().jpb_maxpool_fwd(ptr(x), ptr(y), ptr(idx), B, H, W, Cc, k, s, p, stream_of(x))),
              "jpb_maxpool_fwd")
        ctx.save_for_backward(idx)
        ctx.geom = (B, H, W, Cc, k, s, p)
        return y

    @staticmethod
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        B, H, W, Cc, k, s, p = ctx.geom
        gy = gy if gy.is_contiguous(memory_format=torch.channels_last) else gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty((B, Cc, H, W), dtype=torch.float32, device=gy.device, memory_format=torch.channels_last)   # written once per element
        check(_launch("maxpool_bwd", gy, lambda: _lib.lib().jpb_maxpool_bwd(ptr(gy), ptr(idx), ptr(gx), B, H, W, Cc, k, s, p, stream_of(gy))),
              "jpb_maxpool_bwd")
        return gx, None, None, None


def maxpool(x, k, s, p):
    """nn.MaxPool2d(k, s, p) on a channels-last tensor (C % 4 == 0)."""
    return _MaxPool.apply(x, k, s, p)


# --------------------------------------------------------------------------------------------------
# BatchNorm2d (+ residual) (+ ReLU), NHWC
# --------------------------------------------------------------------------------------------------
def _clast(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


_BN_WS: dict = {}
# Set by TrainEngine.step: parameter gradients are ADDED straight into ``param.grad`` (views of the flat gradient buffer)
# by the backward kernels and autograd receives None for them — no per-parameter AccumulateGrad add / zero-fill launches.
DIRECT_GRAD = False


GRAD_EVENT = None   # set by TrainEngine (world > 1): called with every parameter whose gradient a backward launch is about to write


def direct_grad_target(p):
    """``p.grad`` when the backward kernels may accumulate into it directly (engine step, dense fp32 gradient present)."""
    if not DIRECT_GRAD or p is None:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or g.shape != p.shape or g.stride() != p.stride():
        return None
    if GRAD_EVENT is not None:
        GRAD_EVENT(p)
    return g


def _bn_workspace(C, device):
    """One zero-initialised workspace per (device, stream): the BatchNorm launches of a stream are ordered, launches of
    different streams (the branch-concurrent forward/backward of ``Baseline``) must not share accumulators."""
    need = _lib.lib().jpb_bn_workspace_doubles(max(C, 512))
    key = (device, torch.cuda.current_stream(device).cuda_stream) if device.type == "cuda" else (device, 0)
    ws = _BN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _BN_WS[key] = torch.zeros(need, dtype=torch.float64, device=device)
    return ws


class _BNTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, running_mean, running_var, momentum, eps, relu, nbt, nbt_inc, stats_ready=False):
        x = _clast(x)
        res = _clast(res) if res is not None else None
        B, Cc, H, W = x.shape
        rows = B * H * W
        y = torch.empty_like(x, memory_format=torch.channels_last)
        stat = torch.empty(2 * Cc, dtype=torch.float32, device=x.device)
        ws = _bn_workspace(Cc, x.device)
        check(_launch("bn_fwd", x, lambda: _lib.lib().jpb_bn_train_fwd(
            ptr(x), ptr(res), ptr(gamma.detach()), ptr(beta.detach()), ptr(running_mean), ptr(running_var),
            ptr(nbt) if nbt is not None else None, int(nbt_inc), float(momentum), float(eps),
            int(relu), ptr(y), ptr(stat), ptr(ws), rows, Cc, int(stats_ready), stream_of(x))), "jpb_bn_train_fwd")
        # BatchNorm + ReLU without residual: the backward re-derives the mask y > 0 from x (bit-exact: csrc/bn.cu bn_affine) instead of
        # reading y back in both of its passes
        ctx.save_for_backward(x, y if (relu and res is not None) else None, stat, gamma, beta)
        ctx.cfg = (rows, Cc, int(relu), res is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, stat, gamma, beta = ctx.saved_tensors
        rows, Cc, relu, has_res = ctx.cfg
        gy = _clast(gy)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        dres = torch.empty_like(x, memory_format=torch.channels_last) if (has_res and relu) else None
        tg, tb = direct_grad_target(gamma), direct_grad_target(beta)
        direct = tg is not None and tb is not None
        dgamma = tg if direct else torch.empty(Cc, dtype=torch.float32, device=x.device)
        dbeta = tb if direct else torch.empty(Cc, dtype=torch.float32, device=x.device)
        ws = _bn_workspace(Cc, x.device)
        check(_launch("bn_bwd", x, lambda: _lib.lib().jpb_bn_train_bwd(
            ptr(x), ptr(gy), ptr(y), ptr(stat), ptr(gamma.detach()), relu, ptr(dx), ptr(dres), ptr(dgamma), ptr(dbeta), int(direct),
            ptr(ws), rows, Cc, stream_of(x), ptr(beta.detach()))), "jpb_bn_train_bwd")
        if has_res and not relu:
            dres = gy
        if direct:
            return dx, dres, None, None, None, None, None, None, None, None, None, None
        return dx, dres, dgamma, dbeta, None, None, None, None, None, None, None, None


def batchnorm_train(x, res, gamma, beta, running_mean, running_var, momentum, eps, relu, num_batches_tracked=None, nbt_inc=0,
                    stats_ready=False):
    """``num_batches_tracked`` (int64 [1] device tensor) is advanced by ``nbt_inc`` inside the statistics kernel.
    ``stats_ready``: the convolution that produced ``x`` already accumulated its column sums (``bn_stats_pointer``)."""
    return _BNTrain.apply(x, res, gamma, beta, running_mean, running_var, momentum, eps, relu, num_batches_tracked, nbt_inc, stats_ready)


def bn_stats_pointer(C, device):
    """Device address of the statistics accumulators of the shared BatchNorm workspace (for JpbConvArgs.stats)."""
    ws = _bn_workspace(C, device)
    return _lib.lib().jpb_bn_stats_accumulator(ptr(ws))


def batchnorm_eval(x, res, gamma, beta, running_mean, running_var, eps, relu):
    x = _clast(x)
    res = _clast(res) if res is not None else None
    B, Cc, H, W = x.shape
    stat = torch.cat([running_mean, torch.rsqrt(running_var + eps)]).float().contiguous()
    y = torch.empty_like(x, memory_format=torch.channels_last)
    check(_launch("bn_eval", x, lambda: _lib.lib().jpb_bn_eval_fwd(ptr(x), ptr(res), ptr(gamma.detach()), ptr(beta.detach()), ptr(stat),
                                                                   int(relu), ptr(y), B * H * W, Cc, stream_of(x))), "jpb_bn_eval_fwd")
    return y
