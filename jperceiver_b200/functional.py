"""Autograd bindings of the C-ABI kernels (``include/jpb200.h``): one ``torch.autograd.Function`` per
fused operator.  PyTorch is used for device memory, streams and the autograd tape only; all arithmetic
happens in ``libjpb200.so``."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import PhotoArgs, PhotoGrad, check, ptr, stream_of


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
def _photo_args(disp, target, sources, Ts, noises, K, invK, automask, min_depth, max_depth, noise_scale, seed, stream_id):
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
                        cfg["max_depth"], cfg["noise_scale"], cfg["seed"], cfg["stream"])
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        winner = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
        a.loss_sum, a.winner = ptr(acc), ptr(winner)
        min_index, warped = None, []
        if cfg["debug_outputs"]:
            min_index = torch.empty(B, H, W, dtype=torch.int64, device=dev)
            a.min_index = ptr(min_index)
            for i in range(F):
                w = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
                warped.append(w)
                a.warped[i] = ptr(w)
        check(_lib.lib().jpb_photometric_fwd(C.byref(a), stream_of(target)), "jpb_photometric_fwd")
        loss = finalize(acc, 1.0 / (B * H * W * cfg["num_scales"])).reshape(())
        ctx.cfg = cfg
        ctx.noises = noises
        ctx.save_for_backward(disp_c, K, invK, target, winner, *sources, *Ts)
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
        g = PhotoGrad()
        gl = _f32c(gloss).reshape(1)
        gdisp = torch.zeros_like(disp)
        gT = [torch.zeros_like(T) for T in Ts]
        g.grad_out, g.inv_count, g.winner, g.grad_disp = ptr(gl), 1.0 / (B * H * W * cfg["num_scales"]), ptr(winner), ptr(gdisp)
        for i in range(F):
            g.grad_T[i] = ptr(gT[i])
        check(_lib.lib().jpb_photometric_bwd(C.byref(a), C.byref(g), stream_of(target)), "jpb_photometric_bwd")
        n_extra = len(ctx.noises) if ctx.noises is not None else 0
        return (gdisp, None, None, None, None) + (None,) * F + tuple(gT) + (None,) * n_extra


def photometric_loss(disp, target, sources, Ts, K, invK, *, num_scales=4, automask=True, min_depth=0.1,
                     max_depth=100.0, noise=None, noise_scale=1e-5, seed=0, stream=0, debug_outputs=False):
    """``loss_dict[("min_reconstruct_loss", s)]`` of one scale (already divided by ``num_scales``).

    Returns ``(loss, winner_u8, min_index|None, [warped...])``.  ``noise``: list of B×H×W tensors for the
    identity terms (tests) or None for the in-kernel Philox draw scaled by ``noise_scale``.
    """
    cfg = dict(F=len(sources), num_scales=num_scales, automask=automask, min_depth=min_depth, max_depth=max_depth,
               noise_scale=noise_scale, seed=seed, stream=stream, debug_outputs=debug_outputs)
    rest = list(sources) + list(Ts) + (list(noise) if noise is not None else [])
    out = _Photometric.apply(disp, K, invK, target, cfg, *rest)
    loss, winner = out[0], out[1]
    if debug_outputs:
        return loss, winner, out[2], list(out[3:])
    return loss, winner, None, []
