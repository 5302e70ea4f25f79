"""``MonoDataset.preprocess`` on the device (mono/datasets/mono_dataset.py:126-171, 202-203, 337-343, 417-431).

The reference decodes, flips, resizes twice with ``Image.ANTIALIAS`` (Lanczos), colour-jitters and converts every frame on
dataloader worker processes with PIL, then copies float tensors to the GPU.  Here the decoded uint8 frames of a batch are
copied once and everything after the decoder runs in ``libjpb200.so`` (``csrc/imgpipe.cu``), bit-exact with Pillow's own
arithmetic.  Host work that remains: the JPEG/PNG decode itself, the random draws (same torch RNG calls, same order as
``transforms.ColorJitter``) and the small fixed-point coefficient tables (built once per size pair exactly as Pillow's
``precompute_coeffs`` / ``normalize_coeffs_8bpc``)."""
from __future__ import annotations

import ctypes as C
import functools
import math

import torch

from .._lib import JitterArgs, ResizeArgs, check, lib, ptr, stream_of

PRECISION_BITS = 32 - 8 - 2            # Resample.c
LANCZOS_SUPPORT = 3.0


def _lanczos(x):
    """Resample.c lanczos_filter / sinc_filter (libm ``sin`` through ``math.sin``, as the C code)."""
    if -3.0 <= x < 3.0:
        def sinc(v):
            if v == 0.0:
                return 1.0
            v = v * math.pi
            return math.sin(v) / v
        return sinc(x) * sinc(x / 3)
    return 0.0


@functools.lru_cache(maxsize=64)
def _lanczos_tables_cpu(in_size, out_size):
    """Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc -> (int32 [out,ks], int32 [out,2], ks)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = torch.zeros(out_size, ksize, dtype=torch.int32)
    bounds = torch.zeros(out_size, 2, dtype=torch.int32)
    ss = 1.0 / filterscale
    one = float(1 << PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return kk, bounds, ksize


_TABLE_CACHE = {}


def lanczos_tables(in_size, out_size, device):
    """Device copies of the coefficient tables of one axis; ``(None, None, 0)`` when the size does not change (Pillow skips
    the pass)."""
    if in_size == out_size:
        return None, None, 0
    key = ("lanczos", in_size, out_size, str(device))
    if key not in _TABLE_CACHE:
        kk, bounds, ks = _lanczos_tables_cpu(in_size, out_size)
        _TABLE_CACHE[key] = (kk.to(device), bounds.to(device), ks)
    return _TABLE_CACHE[key]


def nearest_table(in_size, out_size, device):
    """Geometry.c ImagingScaleAffine: source position of every output position, accumulated in double."""
    key = ("nearest", in_size, out_size, str(device))
    if key not in _TABLE_CACHE:
        a = in_size / out_size
        xo = a * 0.5
        tab = []
        for _ in range(out_size):
            tab.append(min(int(xo), in_size - 1))
            xo += a
        _TABLE_CACHE[key] = torch.tensor(tab, dtype=torch.int32, device=device)
    return _TABLE_CACHE[key]


def _u8(t, name):
    if t.dtype != torch.uint8:
        raise TypeError("%s must be uint8 (decoded image bytes)" % name)
    return t.contiguous()


def _flags(flags, B, device):
    if flags is None:
        return None
    f = torch.as_tensor(flags, device=device).to(torch.uint8).contiguous()
    if f.numel() != B:
        raise ValueError("one flag per sample expected")
    return f


def resize_lanczos(src, size, flip=None, want_u8=True, want_float=True):
    """``transforms.Resize(size, interpolation=Image.ANTIALIAS)`` (+ ``ToTensor``) of a batch of decoded frames.

    ``src``: uint8 B×H×W×3 (HWC) on the device; ``size`` = (height, width); ``flip``: per-sample booleans — the horizontal flip
    ``get_color`` applies before resizing.  Returns ``(uint8 B×h×w×3 | None, float32 B×3×h×w | None)``."""
    src = _u8(src, "src")
    B, Hin, Win, ch = src.shape
    if ch != 3:
        raise ValueError("RGB frames expected")
    Hout, Wout = int(size[0]), int(size[1])
    dev = src.device
    a = ResizeArgs()
    kx, bx, a.ksx = lanczos_tables(Win, Wout, dev)
    ky, by, a.ksy = lanczos_tables(Hin, Hout, dev)
    tmp = torch.empty(B, Hin, Wout, 3, dtype=torch.uint8, device=dev)
    out = torch.empty(B, Hout, Wout, 3, dtype=torch.uint8, device=dev) if want_u8 else None
    outf = torch.empty(B, 3, Hout, Wout, dtype=torch.float32, device=dev) if want_float else None
    fl = _flags(flip, B, dev)
    a.src, a.tmp, a.dst, a.dst_f = ptr(src), ptr(tmp), ptr(out), ptr(outf)
    a.B, a.Hin, a.Win, a.Hout, a.Wout = B, Hin, Win, Hout, Wout
    a.kx, a.bx, a.ky, a.by, a.flip = ptr(kx), ptr(bx), ptr(ky), ptr(by), ptr(fl)
    check(lib().jpb_resize_lanczos_u8(C.byref(a), stream_of(src)), "jpb_resize_lanczos_u8")
    return out, outf


def draw_color_jitter(n, brightness=(0.8, 1.2), contrast=(0.8, 1.2), saturation=(0.8, 1.2), hue=(-0.1, 0.1)):
    """``n`` consecutive ``transforms.ColorJitter.get_params`` draws from torch's global RNG, in torchvision's call order
    (randperm(4), then one uniform each for brightness, contrast, saturation, hue) — the same stream a ``ColorJitter`` module
    consumes when it is called ``n`` times.  Returns CPU tensors ``order`` int32 n×4, ``factor`` float32 n×4 (b, c, s, 0) and
    ``hue_shift`` int32 n (``uint8(int32(hue * 255))``), plus the raw float factors for reference."""
    order = torch.zeros(n, 4, dtype=torch.int32)
    factor = torch.zeros(n, 4, dtype=torch.float32)
    hue_shift = torch.zeros(n, dtype=torch.int32)
    raw = []
    for i in range(n):
        order[i] = torch.randperm(4).to(torch.int32)
        b = float(torch.empty(1).uniform_(brightness[0], brightness[1]))
        c = float(torch.empty(1).uniform_(contrast[0], contrast[1]))
        s = float(torch.empty(1).uniform_(saturation[0], saturation[1]))
        h = float(torch.empty(1).uniform_(hue[0], hue[1]))
        factor[i, 0], factor[i, 1], factor[i, 2] = b, c, s       # Pillow's blend takes a C float
        hue_shift[i] = int(h * 255) & 255                         # np.int32(h * 255).astype(np.uint8)
        raw.append((b, c, s, h))
    return order, factor, hue_shift, raw


def color_jitter(src, order, factor, hue_shift, enable=None, want_u8=True, want_float=True):
    """``transforms.ColorJitter`` with given parameters (``draw_color_jitter``) on a batch of uint8 HWC frames; ``enable``:
    per-sample ``do_color_aug`` (disabled samples pass through).  Returns ``(uint8 | None, float32 NCHW | None)``."""
    src = _u8(src, "src")
    B, H, W, ch = src.shape
    if ch != 3:
        raise ValueError("RGB frames expected")
    dev = src.device
    order = torch.as_tensor(order).to(dev, torch.int32).contiguous()
    factor = torch.as_tensor(factor).to(dev, torch.float32).contiguous()
    hue_shift = torch.as_tensor(hue_shift).to(dev, torch.int32).contiguous()
    if order.shape != (B, 4) or factor.shape != (B, 4) or hue_shift.shape != (B,):
        raise ValueError("order / factor must be Bx4 and hue_shift B")
    en = _flags(enable, B, dev)
    out = torch.empty_like(src) if want_u8 else None
    outf = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if want_float else None
    lsum = torch.zeros(B, dtype=torch.int64, device=dev)
    a = JitterArgs()
    a.src, a.dst, a.dst_f, a.B, a.H, a.W = ptr(src), ptr(out), ptr(outf), B, H, W
    a.order, a.factor, a.hue_shift, a.enable, a.lsum = ptr(order), ptr(factor), ptr(hue_shift), ptr(en), ptr(lsum)
    check(lib().jpb_color_jitter_u8(C.byref(a), stream_of(src)), "jpb_color_jitter_u8")
    return out, outf


def bev_label(src, size, flip=None):
    """``process_topview_both`` (mono_dataset.py:425-431): nearest resize to ``size``² and ``== 255 -> 1``; equals
    ``process_topview`` (:417-424) for two-level {0, 255} label images (its ``convert("1")`` dither is then the identity).
    ``src``: uint8 B×H×W (the label's L channel); returns float32 B×1×size×size (``ToTensor`` of the float64 map)."""
    src = _u8(src, "src")
    B, Hin, Win = src.shape
    dev = src.device
    out = torch.empty(B, 1, size, size, dtype=torch.float32, device=dev)
    fl = _flags(flip, B, dev)
    check(lib().jpb_bev_label_u8(ptr(src), ptr(out), B, Hin, Win, int(size), ptr(nearest_table(Win, size, dev)),
                                 ptr(nearest_table(Hin, size, dev)), ptr(fl), stream_of(src)), "jpb_bev_label_u8")
    return out


class GpuPreprocess:
    """The colour part of ``MonoDataset.__getitem__`` + ``preprocess`` for a whole batch.

    ``frames``: ``{frame_id: uint8 B×Hs×Ws×3}`` decoded frames of the snippets (one source size per call; KITTI sequences
    with different native sizes go in separate calls).  Produces the dict entries the model consumes:
    ``("color", f, -1)`` (resize_full), ``("color", f, 0)``, ``("color_aug", f, 0)`` — float32 NCHW in [0, 1]."""

    def __init__(self, height, width, full_res=(375, 1242), is_train=True):
        self.height, self.width, self.full_res, self.is_train = height, width, tuple(full_res), is_train

    def draw(self, batch, n_frames):
        """The per-sample random decisions in the reference's order (mono_dataset.py:202-203): ``do_color_aug`` then
        ``do_flip`` from Python's ``random``, then one ColorJitter draw per frame when augmenting (torch RNG)."""
        import random
        do_aug, do_flip, params = [], [], []
        for _ in range(batch):
            do_aug.append(self.is_train and random.random() > 0.5)
            do_flip.append(self.is_train and random.random() > 0.5)
            params.append(draw_color_jitter(n_frames) if do_aug[-1] else None)
        return do_aug, do_flip, params

    def __call__(self, frames, do_aug, do_flip, params):
        out = {}
        fids = list(frames)
        B = frames[fids[0]].shape[0]
        for j, f in enumerate(fids):
            full_u8, full_f = resize_lanczos(frames[f], self.full_res, flip=do_flip)
            u8, col = resize_lanczos(full_u8, (self.height, self.width))
            out[("color", f, -1)], out[("color", f, 0)] = full_f, col
            order = torch.zeros(B, 4, dtype=torch.int32)
            factor = torch.zeros(B, 4, dtype=torch.float32)
            shift = torch.zeros(B, dtype=torch.int32)
            for b in range(B):
                if do_aug[b]:
                    order[b], factor[b], shift[b] = params[b][0][j], params[b][1][j], params[b][2][j]
            out[("color_aug", f, 0)] = color_jitter(u8, order, factor, shift, enable=do_aug, want_u8=False)[1]
        return out
