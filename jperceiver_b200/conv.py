"""Host side of the tcgen05 implicit-GEMM convolution (``csrc/conv_tc.cu``): the per-layer K-chunk table,
weight packing for ragged channel counts, and the autograd binding.

Backward: ``jpb_act_bwd`` (epilogue backward + bias gradient) -> data gradient = the forward kernel run on dz with
flipped/transposed weights and a scatter epilogue that undoes reflection padding / up-sampling / concatenation ->
weight gradient = ``jpb_conv2d_wgrad`` (pixels are the GEMM reduction, MN-major operands, split over CTAs)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr, stream_of

CL = torch.channels_last
ACT = {"none": 0, "relu": 1, "leaky": 2, "sigmoid": 3}
SMALLN = True      # route 3x3 convolutions with <= 2 output channels to the CUDA-core kernels (csrc/conv_smalln.cu)
_TABLES: dict = {}
_ORDERS: dict = {}
import os as _os
# K-block order of the forward / data-gradient GEMM: "cb" = channel-block major (the taps of one 32-channel block are
# consecutive, so their gathers re-read the same pixels and hit L1), "tap" = the natural weight order
KORDER = _os.environ.get("JPB_CONV_KORDER", "cb")
DBG_SKIP = int(_os.environ.get('JPB_CONV_SKIP', '0'))   # timing experiments (wrong results): 1 no A gather, 2 no weight TMA
DBG_STAMPS = None   # int64 [512, 6, 8] device tensor: per-CTA timeline of the next forward launches (tools/conv_timeline.py)
L1_GATHER = int(_os.environ.get("JPB_CONV_L1", "1"))
# Arithmetic of the tensor-core convolutions (same kernels, same path):
#   "tf32"   one TF32 product per term — what the reference gets from cuDNN with torch.backends.cudnn.allow_tf32 = True (torch default)
#   "3xtf32" operands split into TF32 hi + lo parts (jpb_tf32_split), hi*hi + hi*lo + lo*hi accumulated in the same TMEM tile
#            as a 3x longer K — fp32-grade results (reference with allow_tf32 = False / its CPU path), ~3x the tensor work
PRECISION = _os.environ.get("JPB_CONV_PRECISION", "tf32")
# kind::tf32 truncates its operands (cuDNN's TF32 kernels round to nearest); the mean shortfall of a product of two truncated
# operands, 2 * 2^-11 * ln 2, is folded back into the accumulator (include/jpb200.h: JpbConvArgs.acc_scale).  Measured effect on
# the whole network: tests/test_model_parity.py::test_full_size_gpu_tf32_calibrated_against_cudnn_tf32.
TRUNC_COMP = float(_os.environ.get("JPB_TF32_TRUNC_COMP", "1.00067702"))


def _acc_scale():
    return 1.0 if PRECISION == "3xtf32" else TRUNC_COMP


class trunc_comp:
    """``with conv.trunc_comp(1.0): ...`` — scoped override of the truncation compensation (tests that feed operands which ARE
    TF32-representable, so that nothing is truncated and the products are exact)."""

    def __init__(self, value):
        self.value = float(value)

    def __enter__(self):
        global TRUNC_COMP
        self.old, TRUNC_COMP = TRUNC_COMP, self.value
        return self

    def __exit__(self, *exc):
        global TRUNC_COMP
        TRUNC_COMP = self.old


class precision:
    """``with conv.precision("3xtf32"): ...`` — scoped switch of the convolution arithmetic (tests, parity runs)."""

    def __init__(self, mode):
        if mode not in ("tf32", "3xtf32"):
            raise ValueError("conv precision %r (tf32 | 3xtf32)" % (mode,))
        self.mode = mode

    def __enter__(self):
        global PRECISION
        self.old, PRECISION = PRECISION, self.mode
        return self

    def __exit__(self, *exc):
        global PRECISION
        PRECISION = self.old


def tf32_split(x):
    """[B, C, H, W] channels-last (or [rows, C] row-major) -> the hi | lo halves side by side: [B, 2*Cp, H, W] / [rows, 2*Cp]."""
    if x.dim() == 4:
        x = _cl(x)
        B, Cc, H, W = x.shape
        Cp = _pad4(Cc)
        out = torch.empty((B, 2 * Cp, H, W), dtype=torch.float32, device=x.device, memory_format=CL)
        rows = B * H * W
    else:
        x = x.contiguous()
        rows, Cc = x.shape
        Cp = _pad4(Cc)
        out = torch.empty((rows, 2 * Cp), dtype=torch.float32, device=x.device)
    check(_launch("tf32_split", x, lambda: _lib.lib().jpb_tf32_split(ptr(x), ptr(out), rows, Cc, stream_of(x))), "jpb_tf32_split")
    return out


class _WTCache:
    """Flipped/transposed weights for the data-gradient GEMMs, rebuilt for all registered layers by ONE launch per step
    (``refresh``; called by TrainEngine.step while the weights are final for the step).  Outside an engine step the data
    gradient re-lays out its weight per call, as before."""

    def __init__(self):
        self.enabled = False     # inside TrainEngine.step
        self.fresh = False       # buffers hold this step's weights
        self.entries = {}        # (data_ptr, shape) -> (weight, wT)
        self.table = None        # device copy of the descriptor array
        self.nblocks = 0
        self.dirty = False

    def lookup(self, weight):
        if not (self.enabled and self.fresh):
            return None
        e = self.entries.get((weight.data_ptr(), tuple(weight.shape)))
        return e[1] if e is not None else None

    def register(self, weight):
        if not self.enabled:
            return
        key = (weight.data_ptr(), tuple(weight.shape))
        if key not in self.entries:
            N, Cin, kh, kw = weight.shape
            wT = torch.empty((Cin, N, kh, kw), dtype=torch.float32, device=weight.device).contiguous(memory_format=CL)
            self.entries[key] = (weight, wT)
            self.dirty = True

    def refresh(self):
        if not self.entries:
            self.fresh = True
            return
        if self.dirty or self.table is None:
            arr = (_lib.WeightT * len(self.entries))()
            start = 0
            for i, (w, wT) in enumerate(self.entries.values()):
                N, Cin, kh, kw = w.shape
                arr[i].src, arr[i].dst, arr[i].N, arr[i].Cin, arr[i].taps, arr[i].block_start = ptr(w), ptr(wT), N, Cin, kh * kw, start
                start += kh * kw * ((N + 31) // 32) * ((Cin + 31) // 32)
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
            dev = next(iter(self.entries.values()))[0].device
            self.table = raw.to(dev)
            self.nblocks = start
            self.dirty = False
        t = self.table
        check(_launch("weight_flipT", t, lambda: _lib.lib().jpb_weight_flipT(ptr(t), len(self.entries), self.nblocks, stream_of(t))),
              "jpb_weight_flipT")
        self.fresh = True


WT = _WTCache()


def _pad4(c):
    return (c + 3) // 4 * 4


def chunk_table(src_channels, kh, kw, device):
    """int32 [nkb*8, 4] table: one row per 16-byte chunk of K in weight order (tap-major, then source, then channel)."""
    key = (tuple(src_channels), kh, kw, str(device))
    t = _TABLES.get(key)
    if t is None:
        rows = []
        nsrc = len(src_channels)
        for ky in range(kh):
            for kx in range(kw):
                tap = ky * kw + kx
                for si, c in enumerate(src_channels):
                    for coff in range(0, c, 4):
                        rows.append((si | ((tap * nsrc + si) << 8), (ky << 16) | (kx & 0xFFFF), coff, min(16, (c - coff) * 4)))
        while len(rows) % 8:
            rows.append((-1, 0, 0, 0))
        t = _TABLES[key] = torch.tensor(rows, dtype=torch.int32, device=device).contiguous()
    return t


def split_table(src_cp, kh, kw, device, part):
    """Chunk table over split sources (``tf32_split``: pixel stride 2*Cp, hi half at channel 0, lo half at channel Cp).
    ``part``: "hi" / "lo" = one half (weight-gradient launches); "3x" = the K extension hi | hi | lo that pairs with the weight
    matrix [W_hi | W_lo | W_hi] (forward and data gradient)."""
    key = ("split", part, tuple(src_cp), kh, kw, str(device))
    t = _TABLES.get(key)
    if t is None:
        hi, lo = [], []
        nsrc = len(src_cp)
        for ky in range(kh):
            for kx in range(kw):
                tap = ky * kw + kx
                for si, c in enumerate(src_cp):
                    for coff in range(0, c, 4):
                        hi.append((si | ((tap * nsrc + si) << 8), (ky << 16) | (kx & 0xFFFF), coff, 16))
                        lo.append((si | ((tap * nsrc + si) << 8), (ky << 16) | (kx & 0xFFFF), c + coff, 16))
        rows = {"hi": hi, "lo": lo, "3x": hi + hi + lo}[part]
        while len(rows) % 8:
            rows.append((-1, 0, 0, 0))
        t = _TABLES[key] = torch.tensor(rows, dtype=torch.int32, device=device).contiguous()
    return t


def ordered_table(src_channels, kh, kw, device, split3=False):
    """(table, kcol): the chunk table with its K blocks (8 rows each) re-ordered channel-block major, and the weight column
    of each K block (int32 [nkb]) for the weight TMA.  Any order is valid — K is the GEMM reduction."""
    base = (lambda: split_table(src_channels, kh, kw, device, "3x")) if split3 else (lambda: chunk_table(src_channels, kh, kw, device))
    if KORDER != "cb" or kh * kw == 1:
        return base(), None
    key = (tuple(src_channels), kh, kw, str(device), bool(split3))
    r = _ORDERS.get(key)
    if r is None:
        t = base().cpu().view(-1, 8, 4)
        first = t[:, 0, :]                                       # first chunk of each K block: (src | tapsrc<<8, dydx, coff, bytes)
        nsrc = len(src_channels)
        keys = []
        for i in range(t.shape[0]):
            x, _, coff, _ = [int(v) for v in first[i]]
            if x < 0:
                keys.append((1 << 30, 0, 0, i))
            else:
                si, tap = x & 0xff, (x >> 8) // nsrc
                keys.append((si, coff // 32, tap, i))
        perm = torch.tensor([k[3] for k in sorted(keys)], dtype=torch.long)
        tab = t[perm].reshape(-1, 4).contiguous().to(device)
        kcol = (perm * 32).to(torch.int32).to(device)
        r = _ORDERS[key] = (tab, kcol)
    return r


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


def gemm_weight3(weight, src_channels, weight_channels):
    """3xTF32 operand [N, 3*K1] = [W_hi | W_lo | W_hi] over the zero-padded packed K layout (every source padded to 4 channels),
    matching ``split_table(..., "3x")``."""
    N, Cin, kh, kw = weight.shape
    w = weight.permute(0, 2, 3, 1)
    parts, off = [], 0
    for c_t, c_w in zip(src_channels, weight_channels):
        parts.append(F.pad(w[..., off:off + c_w], (0, _pad4(c_t) - c_w)))
        off += c_w
    wp = (parts[0] if len(parts) == 1 else torch.cat(parts, -1)).contiguous().reshape(N, -1)
    K1 = wp.shape[1]
    hl = tf32_split(wp)                         # [N, 2*K1]: hi | lo
    w3 = torch.cat([hl, hl[:, :K1]], 1).contiguous()
    return w3, 3 * K1


SPLIT_EPILOGUE = int(_os.environ.get('JPB_SPLIT_EPILOGUE', '1'))   # split-K also for convolutions with bias / activation (finishing pass)
KSPLIT_SLOTS = int(_os.environ.get("JPB_KSPLIT_SLOTS", "296"))


def _ksplit(M, N, nkb):
    """Split-K factor for tiles that cannot fill the machine (deep, small-extent layers: layer3/4, pose and layout tails).
    Target: two shallow CTAs per SM (the wide-tile kernel then runs its 2-CTA/SM configuration), >= 4 K blocks per CTA."""
    nt = 16
    while nt < N and nt < 256:
        nt *= 2
    tiles = ((M + 127) // 128) * ((N + nt - 1) // nt)
    if tiles >= 74 or nkb < 16 or KSPLIT_SLOTS <= 0:
        return 1
    if KSPLIT_SLOTS <= 148:
        return max(1, min(148 // tiles, nkb // 8, 16))
    return max(1, min(KSPLIT_SLOTS // tiles, nkb // 4, 32))


def _fill_sources(a, xs, ups):
    for i, x in enumerate(xs):
        a.src[i] = ptr(x)
        a.src_C[i], a.src_H[i], a.src_W[i], a.src_up[i] = x.shape[1], x.shape[2], x.shape[3], int(ups[i])
    a.nsrc = len(xs)


def _cl(x):
    return x if x.is_contiguous(memory_format=CL) else x.contiguous(memory_format=CL)


def _launch(name, t, call, tag=None):
    from .functional import _launch as L
    return L(name, t, call, tag)


def act_bwd(gy, y, act, want_bias, bias_target=None):
    """dz = gy * act'(y) [+ bias gradient] in one pass over NHWC data.  ``bias_target``: accumulate the bias gradient
    straight into this tensor (the parameter's slot of the flat gradient buffer) instead of a fresh zero-filled one."""
    gy = _cl(gy)
    B, N, Ho, Wo = gy.shape
    rows = B * Ho * Wo
    need_dz = act != "none"
    if not need_dz and not want_bias:
        return gy, None
    dz = torch.empty_like(gy, memory_format=CL) if need_dz else None
    gb = (bias_target if bias_target is not None else torch.zeros(N, dtype=torch.float32, device=gy.device)) if want_bias else None
    check(_launch("act_bwd", gy, lambda: _lib.lib().jpb_act_bwd(ptr(gy), ptr(y) if need_dz else None, ptr(dz), rows, N, ACT[act],
                                                                 ptr(gb), stream_of(gy))), "jpb_act_bwd")
    return (dz if need_dz else gy), gb


S2_CLASSES = int(_os.environ.get("JPB_DGRAD_S2_CLASSES", "1"))   # 0: zero-stuffed stride-2 data gradient (in_div = 2) everywhere
S2_MIN_PIXELS = int(_os.environ.get("JPB_DGRAD_S2_MIN_PIXELS", "16384"))   # below: four latency-sized launches cost more than the 27 zero taps
_S2_TABLES: dict = {}


def _s2_class_table(Nc, cy, cx, device):
    """(table, kcol, kh_c, kw_c) of parity class (cy, cx) of the data gradient of a 3x3 / stride-2 / pad-1 convolution: input-gradient
    pixels (2i + cy, 2j + cx) only see the flipped taps ky' in ({1}, {0, 2})[cy], kx' likewise; as a (1 + cy) x (1 + cx) stride-1
    convolution over dz reading dz[i + ty, j + tx] (zero outside).  ``kcol``: column of each K block in the full flipped /
    transposed weight matrix [Cin][9 * Nc]."""
    key = (Nc, cy, cx, str(device))
    r = _S2_TABLES.get(key)
    if r is None:
        kh_c, kw_c = 1 + cy, 1 + cx
        rows, kcol = [], []
        for ty in range(kh_c):
            for tx in range(kw_c):
                t = ty * kw_c + tx
                kyf, kxf = (2 * ty if cy else 1), (2 * tx if cx else 1)
                for cb in range(0, Nc, 32):
                    kcol.append((kyf * 3 + kxf) * Nc + cb)
                    for coff in range(cb, cb + 32, 4):
                        rows.append((0 | (t << 8), (ty << 16) | tx, coff, 16))
        r = _S2_TABLES[key] = (torch.tensor(rows, dtype=torch.int32, device=device).contiguous(),
                               torch.tensor(kcol, dtype=torch.int32, device=device).contiguous(), kh_c, kw_c)
    return r


def _dgrad_s2_classes(dz, wmat, wcols, Cin, x, B, Ho, Wo, H, W):
    """Data gradient of a 3x3 / stride-2 / pad-1 convolution as four parity-class launches (9 taps in all instead of 36 zero-stuffed
    ones); every element of the gradient is written exactly once."""
    Nc = dz.shape[1]
    dev = dz.device
    grad = torch.empty_like(x, memory_format=CL)
    for cy in (0, 1):
        for cx in (0, 1):
            table, kcol, kh_c, kw_c = _s2_class_table(Nc, cy, cx, dev)
            a = _lib.ConvArgs()
            a.src[0] = ptr(dz)
            a.src_C[0], a.src_H[0], a.src_W[0], a.src_up[0] = Nc, Ho, Wo, 0
            a.nsrc = 1
            a.B, a.Hin, a.Win = B, Ho, Wo
            a.Ho, a.Wo, a.N = H // 2, W // 2, Cin
            a.stride, a.pad, a.reflect = 1, 0, 0
            a.weight, a.w_row, a.w_cols = ptr(wmat), wmat.stride(0), wcols
            a.table, a.nkb = ptr(table), table.shape[0] // 8
            a.ntaps, a.kw = kh_c * kw_c, kw_c
            a.kcol = ptr(kcol)
            a.acc_scale = _acc_scale()
            a.scatter, a.ndst = 1, 1
            a.dst[0] = ptr(grad)
            a.dst_C[0], a.dst_H[0], a.dst_W[0], a.dst_up[0] = Cin, H, W, 0
            a.fold_H, a.fold_W = H // 2, W // 2
            a.dst_mul, a.dst_oy, a.dst_ox = 2, cy, cx
            if FWD_ROWS and Nc % 32 == 0 and Cin > 32 and (W // 2) % 32 == 0:
                a.rows, a.rows_wv = FWD_ROWS, W // 2
            tag = (B * (H // 2) * (W // 2), Cin, table.shape[0] * 4, 3, 2, (Cin,), (0,), 0, 1)
            check(_launch("conv_dgrad", dz, lambda: _lib.lib().jpb_conv2d_fwd(C.byref(a), stream_of(dz)), tag), "jpb_conv2d_fwd(dgrad, stride-2 class)")
    return [grad]


def conv_dgrad(dz, weight, xs, ups, stride, pad, reflect, needs):
    """Gradients w.r.t. the forward sources, through the forward kernel run on dz with the flipped/transposed weights."""
    N, Cin, kh, kw = weight.shape
    B, _, Ho, Wo = dz.shape
    H = xs[0].shape[2] * (2 if ups[0] else 1)
    W = xs[0].shape[3] * (2 if ups[0] else 1)
    dev = dz.device
    Nc = dz.shape[1]                                   # possibly channel-padded dz
    three = PRECISION == "3xtf32"
    wT = WT.lookup(weight) if (Nc == N and N % 4 == 0 and weight.is_contiguous(memory_format=CL)) else None
    if wT is None:
        wT = weight.detach().flip(2, 3).permute(1, 0, 2, 3).contiguous(memory_format=CL)   # [Cin][kh][kw][Cout]
        if Nc == N and N % 4 == 0 and weight.is_contiguous(memory_format=CL):
            WT.register(weight)
    if (S2_CLASSES and not three and stride == 2 and kh == 3 and kw == 3 and pad == 1 and not reflect and len(xs) == 1 and not ups[0]
            and xs[0].shape[1] == Cin and Nc % 32 == 0 and Cin % 4 == 0 and H % 2 == 0 and W % 2 == 0 and Ho == H // 2 and Wo == W // 2
            and B * H * W >= S2_MIN_PIXELS):
        wmat_c, wcols_c = gemm_weight(wT, [Nc], [N])
        if wcols_c == 9 * Nc:
            return _dgrad_s2_classes(dz, wmat_c, wcols_c, Cin, xs[0], B, Ho, Wo, H, W)
    if three:
        dz_k = tf32_split(dz)                           # [B, 2*Nc, Ho, Wo]: hi | lo halves of every pixel
        wmat, wcols = gemm_weight3(wT, [Nc], [N])
        table, kcol = ordered_table([Nc], kh, kw, dev, split3=True)
    else:
        dz_k = dz
        wmat, wcols = gemm_weight(wT, [Nc], [N])
        table, kcol = ordered_table([Nc], kh, kw, dev)
    src_C = [x.shape[1] for x in xs]
    if len(xs) == 1 and src_C[0] != Cin:                # zero-padded stem input: gradient not needed (images)
        return [None]
    grads = []
    simple = len(xs) == 1 and not ups[0] and not reflect
    for x, up in zip(xs, ups):
        grads.append(torch.empty_like(x, memory_format=CL) if simple else torch.zeros_like(x, memory_format=CL))
    a = _lib.ConvArgs()
    a.src[0] = ptr(dz_k)
    a.src_C[0], a.src_H[0], a.src_W[0], a.src_up[0] = dz_k.shape[1], Ho, Wo, 0
    a.nsrc = 1
    a.B, a.Hin, a.Win = B, Ho, Wo
    a.N = Cin
    a.stride, a.reflect = 1, 0
    a.in_div = 2 if stride == 2 else 0
    assert stride in (1, 2)
    if reflect:
        a.Ho, a.Wo, a.pad = H + 2 * pad, W + 2 * pad, kh - 1
        a.fold_pad, a.fold_reflect = pad, 1
    else:
        a.Ho, a.Wo, a.pad = H, W, kh - 1 - pad
        a.fold_pad, a.fold_reflect = 0, 0
    a.fold_H, a.fold_W = H, W
    a.weight, a.w_row, a.w_cols = ptr(wmat), wmat.stride(0), wcols
    a.table, a.nkb = ptr(table), table.shape[0] // 8
    a.ntaps, a.kw = kh * kw, kw
    a.kcol, a.l1_gather = (ptr(kcol) if kcol is not None else None), int(L1_GATHER and kcol is not None)
    a.act = 0
    a.acc_scale = _acc_scale()
    ks = _ksplit(B * a.Ho * a.Wo, Cin, table.shape[0] // 8)
    if ks > 1:
        a.ksplit = ks
        if simple:
            grads[0].zero_()
    if simple:
        a.out = ptr(grads[0])
        if (not three) and wcols == kh * kw * Nc and _patch_ok(Nc, Cin, H, W, B, kh, kw, stride, pad, False, 1, False, ks) and a.pad == 1:
            a.patch, a.patch_desc_mode = 1 + PATCH_TILE_ROWS, PATCH_DESC_MODE
            _patch_taps(a, TAPS_3X3, 2, (1, 1))
    else:
        a.scatter = 1
        a.ndst = len(xs)
        nt = 256
        for g, up in zip(grads, ups):
            pass
        for i, (g, up) in enumerate(zip(grads, ups)):
            a.dst[i] = ptr(g)
            a.dst_C[i], a.dst_H[i], a.dst_W[i], a.dst_up[i] = g.shape[1], g.shape[2], g.shape[3], int(up)
            if i + 1 < len(grads):
                while g.shape[1] % nt:
                    nt //= 2
        if len(grads) == 1:
            nt = 0
        assert nt == 0 or nt >= 16
        a.nt = nt
    if (FWD_ROWS and not three and not a.patch and stride == 1 and Nc % 32 == 0 and Cin > 32 and (a.nt == 0 or a.nt >= 64)):
        # TMA-row operand (dz is one dense source): the tile raster's row length is the gradient domain's, rounded up to 32 pixels
        # (reflection-padded layers: W + 2 -> the surplus pixels are computed and dropped) when that costs at most half as much again (measured in the step: 125 / 150 / 200 percent -> 18.99 / 18.79 / 19.05 ms)
        wv = (a.Wo + 31) // 32 * 32
        if wv * 100 <= a.Wo * ROWS_WV_MAX:
            a.rows, a.rows_wv = FWD_ROWS, wv
    tag = (B * a.Ho * a.Wo, Cin, table.shape[0] * 4, kh, stride, tuple(src_C), tuple(int(u) for u in ups), int(bool(reflect)), ks)
    check(_launch("conv_dgrad", dz, lambda: _lib.lib().jpb_conv2d_fwd(C.byref(a), stream_of(dz)), tag), "jpb_conv2d_fwd(dgrad)")
    return grads


WGRAD_SLOTS = int(_os.environ.get("JPB_WGRAD_SLOTS", "222"))   # CTAs a weight-gradient launch is split into (pixel ranges x K tiles x N tiles)
WGRAD_ROWS = int(_os.environ.get("JPB_WGRAD_ROWS", "1"))     # 0: always the gathered operand (A/B measurements); 2: deep CTAs for N tile 256
_ROW_TABLES: dict = {}


def upsample2x(x):
    """Nearest 2x up-sampling of a channels-last map (csrc/elementwise.cu: upsample2x_kernel)."""
    x = _cl(x)
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc, 2 * H, 2 * W), dtype=torch.float32, device=x.device, memory_format=CL)
    check(_launch("upsample2x", x, lambda: _lib.lib().jpb_upsample2x(ptr(x), ptr(y), B, H, W, Cc, stream_of(x))), "jpb_upsample2x")
    return y


ROWS_WV_MAX = int(_os.environ.get("JPB_ROWS_WV_MAX", "150"))   # padded raster of a data gradient: at most this many percent of the real row length
FWD_ROWS = int(_os.environ.get("JPB_FWD_ROWS", "1"))         # 0: gathered A operand in the forward / data-gradient kernels (A/B measurements)


def _fwd_rows_ok(src_C, ups, N, kh, kw, stride, pad, reflect, Hin, Win, Ho, Wo):
    """TMA-row A operand of conv_tc_fwd_kernel (JpbConvArgs.rows): stride-1 same-size convolutions with rows of a multiple of 32
    pixels; every source becomes a dense tensor of whole 32-channel blocks (up-sampled ones are materialised, a narrow one padded)."""
    return (FWD_ROWS and PRECISION == "tf32" and stride == 1 and kh == kw and pad == (kh - 1) // 2 and Ho == Hin and Wo == Win and Wo % 32 == 0
            and N > 32 and (not reflect or (pad == 1 and Win >= 64)) and all(c % 32 == 0 or (c < 32 and not u) for c, u in zip(src_C, ups)))


def pad_channels(x, Cp):
    """Zero-padded channel copy of a channels-last map (csrc/elementwise.cu: pad_channels_kernel)."""
    x = _cl(x)
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cp, H, W), dtype=torch.float32, device=x.device, memory_format=CL)
    check(_launch("pad_channels", x, lambda: _lib.lib().jpb_pad_channels(ptr(x), ptr(y), B * H * W, Cc, Cp, stream_of(x))), "jpb_pad_channels")
    return y


def _wgrad_rows_ok(src_C, ups, kh, kw, stride, pad, Hin, Win, Ho, Wo, reflect):
    return (WGRAD_ROWS and PRECISION == "tf32" and stride == 1 and kh == kw and kh in (1, 3) and pad == (kh - 1) // 2 and Ho == Hin and Wo == Win
            and Wo % 32 == 0 and (not reflect or Win >= 64) and src_C[0] % 32 == 0
            and all(c % 32 == 0 or (c < 32 and not u) for c, u in zip(src_C, ups)))


def row_table(src_C, kh, kw, device, tma_C=None):
    """(table, gflags, chunk_col) of the TMA-row weight gradient: first one group of 8 table rows per (tap, source with C % 32 == 0,
    32-channel block) — read as one tensor box —, then the 16-byte chunks of the remaining sources (gathered), in the table format
    of ``chunk_table``.  ``chunk_col``: column of each row's first K position in the packed weight layout [kh, kw, sum(pad4(C))].
    ``tma_C``: channels of the tensors handed to the kernel when a narrow source was zero-padded to a 32-channel block
    (``src_C`` keeps the real counts: the padding channels have no column)."""
    tma_C = list(tma_C) if tma_C is not None else list(src_C)
    key = (tuple(src_C), tuple(tma_C), kh, kw, str(device))
    r = _ROW_TABLES.get(key)
    if r is None:
        nsrc = len(src_C)
        cpad = [_pad4(c) for c in src_C]
        soff = [sum(cpad[:i]) for i in range(nsrc)]
        ctot = sum(cpad)
        rows, cols, flags = [], [], []
        for ky in range(kh):
            for kx in range(kw):
                tap = ky * kw + kx
                for si, c in enumerate(tma_C):
                    if c % 32:
                        continue
                    for cb in range(0, c, 32):
                        for coff in range(cb, cb + 32, 4):
                            rows.append((si | ((tap * nsrc + si) << 8), (ky << 16) | (kx & 0xFFFF), coff, 16))
                            cols.append(tap * ctot + soff[si] + coff if coff < cpad[si] else -1)
                        flags.append(1)
        for ky in range(kh):
            for kx in range(kw):
                tap = ky * kw + kx
                for si, c in enumerate(tma_C):
                    if c % 32 == 0:
                        continue
                    for coff in range(0, c, 4):
                        rows.append((si | ((tap * nsrc + si) << 8), (ky << 16) | (kx & 0xFFFF), coff, min(16, (c - coff) * 4)))
                        cols.append(tap * ctot + soff[si] + coff)
        while len(rows) % 8:
            rows.append((-1, 0, 0, 0))
            cols.append(-1)
        flags += [0] * (len(rows) // 8 - len(flags))
        r = _ROW_TABLES[key] = (torch.tensor(rows, dtype=torch.int32, device=device).contiguous(),
                                torch.tensor(flags, dtype=torch.uint8, device=device).contiguous(),
                                torch.tensor(cols, dtype=torch.int32, device=device).contiguous())
    return r


def conv_wgrad(dz, weight, xs, ups, stride, pad, reflect, dbg=None, target=None, dense=None):
    """``target``: the weight's channels-last gradient view; when given (and the K layout needs no padding) the kernel adds
    into it and None is returned.  ``dense``: the sources as the TMA-row forward already materialised them (up-sampled /
    channel-padded), if it did."""
    N, Cin, kh, kw = weight.shape
    B, Nc, Ho, Wo = dz.shape
    dev = dz.device
    src_C = [x.shape[1] for x in xs]
    w_C = [Cin] if (len(xs) == 1 and src_C[0] != Cin) else src_C
    three = PRECISION == "3xtf32"
    table = chunk_table(src_C, kh, kw, dev)
    raw = all(c % 4 == 0 for c in src_C) and src_C == w_C
    Cpad = sum(_pad4(c) for c in src_C)
    wcols = kh * kw * (Cin if raw else Cpad)
    direct = target is not None and raw and Nc == N and target.permute(0, 2, 3, 1).is_contiguous()
    dw = target if direct else torch.zeros(Nc, wcols, dtype=torch.float32, device=dev)
    a = _lib.ConvWgradArgs()
    a.accumulate = int(direct)
    a.acc_scale = _acc_scale()
    _fill_sources(a, xs, ups)
    a.B = B
    a.Hin = xs[0].shape[2] * (2 if ups[0] else 1)
    a.Win = xs[0].shape[3] * (2 if ups[0] else 1)
    a.Ho, a.Wo, a.N = Ho, Wo, Nc
    a.stride, a.pad, a.reflect = stride, pad, int(reflect)
    a.table, a.nchunks = ptr(table), table.shape[0]
    a.dy, a.dw, a.w_row, a.w_cols = ptr(dz), ptr(dw), wcols, wcols
    if dbg is None and _wgrad_rows_ok(src_C, ups, kh, kw, stride, pad, a.Hin, a.Win, Ho, Wo, reflect):
        # TMA-row operand: dense full-resolution sources (an up-sampled source is materialised), group-major table
        # (and a narrow one — the 1-channel disparity of the iconv layers — zero-padded to one 32-channel block: no gathered group)
        xs_d = dense if dense is not None else [upsample2x(x) if u else (pad_channels(x, 32) if x.shape[1] % 32 else _cl(x)) for x, u in zip(xs, ups)]
        _fill_sources(a, xs_d, [False] * len(xs_d))
        table, gflags, ccol = row_table(src_C, kh, kw, dev, [x.shape[1] for x in xs_d])
        a.table, a.nchunks = ptr(table), table.shape[0]
        a.rows, a.gflags, a.chunk_col = WGRAD_ROWS, ptr(gflags), ptr(ccol)
    nt = 32
    while nt < Nc and nt < 256:
        nt *= 2
    tiles = ((table.shape[0] + 31) // 32) * ((Nc + nt - 1) // nt)
    steps = (B * Ho * Wo + 31) // 32
    a.splits = max(1, min(steps, (WGRAD_SLOTS + tiles - 1) // tiles))
    if dbg is not None:
        a.dbg = ptr(dbg)
    tag = (B * Ho * Wo, Nc, table.shape[0] * 4, kh, stride, tuple(src_C), tuple(int(u) for u in ups), int(bool(reflect)), a.splits)
    if three:
        # dW = im2col(x_hi)^T dz_hi + im2col(x_hi)^T dz_lo + im2col(x_lo)^T dz_hi: three accumulating launches of the same kernel
        src_cp = [_pad4(c) for c in src_C]
        xs_k = [tf32_split(x) for x in xs]
        dz_k = tf32_split(dz)
        _fill_sources(a, xs_k, ups)
        a.accumulate, a.dy_pitch = 1, dz_k.shape[1]
        t_hi, t_lo = split_table(src_cp, kh, kw, dev, "hi"), split_table(src_cp, kh, kw, dev, "lo")
        assert t_hi.shape[0] == table.shape[0]
        for tab, dyoff in ((t_hi, 0), (t_hi, Nc), (t_lo, 0)):
            a.table, a.dy = ptr(tab), ptr(dz_k) + 4 * dyoff
            check(_launch("conv_wgrad", dz, lambda: _lib.lib().jpb_conv2d_wgrad(C.byref(a), stream_of(dz)), tag), "jpb_conv2d_wgrad(3xtf32)")
    else:
        check(_launch("conv_wgrad", dz, lambda: _lib.lib().jpb_conv2d_wgrad(C.byref(a), stream_of(dz)), tag), "jpb_conv2d_wgrad")
    if direct:
        return None
    dw = dw[:N]
    if raw:
        return dw.view(N, kh, kw, Cin).permute(0, 3, 1, 2)
    parts, off = [], 0
    dw = dw.view(N, kh, kw, Cpad)
    for c_t, c_w in zip(src_C, w_C):
        parts.append(dw[..., off:off + c_w])
        off += _pad4(c_t)
    return torch.cat(parts, -1).permute(0, 3, 1, 2)


def _is_smalln(weight, xs, stride, pad, residual):
    N, Cin, kh, kw = weight.shape
    return (N <= 2 and kh == 3 and kw == 3 and stride == 1 and pad == 1 and len(xs) == 1 and residual is None
            and xs[0].shape[1] == Cin and Cin % 4 == 0 and N * 9 * Cin * 4 <= 48 * 1024)


def _w_nhwc(weight):
    w = weight.detach().permute(0, 2, 3, 1)
    return w if w.is_contiguous() else w.contiguous()


def smalln_fwd(x, up, weight, bias, reflect, act):
    B, Cin, Hs, Ws = x.shape
    N = weight.shape[0]
    Ho, Wo = (2 * Hs, 2 * Ws) if up else (Hs, Ws)
    out = torch.empty((B, N, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=CL)
    work = torch.empty(B * Hs * Ws * N * 9, dtype=torch.float32, device=x.device)
    w = _w_nhwc(weight)
    check(_launch("conv_smalln_fwd", x, lambda: _lib.lib().jpb_conv3x3_smalln_fwd(
        ptr(x), ptr(w), ptr(bias.detach()) if bias is not None else None, ptr(out), ptr(work), B, Hs, Ws, Cin, int(up), N, int(reflect),
        ACT[act], stream_of(x))), "jpb_conv3x3_smalln_fwd")
    return out


def smalln_bwd(x, up, dz, weight, reflect, want_dw=True, want_dx=True):
    """(dw, dx) of the small-N convolution from dz (gradient w.r.t. the pre-activation output)."""
    B, Cin, Hs, Ws = x.shape
    N = weight.shape[0]
    work = torch.zeros(B * Hs * Ws * N * 9, dtype=torch.float32, device=x.device)
    dw = torch.zeros(N, 3, 3, Cin, dtype=torch.float32, device=x.device) if want_dw else None
    dx = torch.empty_like(x, memory_format=CL) if want_dx else None
    w = _w_nhwc(weight)
    check(_launch("conv_smalln_bwd", x, lambda: _lib.lib().jpb_conv3x3_smalln_bwd(
        ptr(x), ptr(w), ptr(dz), ptr(work), ptr(dw), ptr(dx), B, Hs, Ws, Cin, int(up), N, int(reflect), stream_of(x))),
        "jpb_conv3x3_smalln_bwd")
    return (dw.permute(0, 3, 1, 2) if want_dw else None), dx


class _ConvTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, weight, bias, residual, *xs):
        ups, stride, pad, reflect, act = cfg["ups"], cfg["stride"], cfg["pad"], cfg["reflect"], cfg["act"]
        xs = [x if x.is_contiguous(memory_format=CL) else x.contiguous(memory_format=CL) for x in xs]
        if SMALLN and _is_smalln(weight, xs, stride, pad, residual):
            STATS_FUSED[0] = False
            out = smalln_fwd(xs[0], ups[0], weight, bias, reflect, act)
            ctx.cfg = cfg
            ctx.save_for_backward(weight, bias, residual, out if act != "none" else None, *xs)
            return out
        if residual is None and _stem_ok(weight, xs, ups, stride, pad, reflect, 1):
            out = stem_forward(xs[0], weight, bias, act, cfg.get("bn_stats"))
            ctx.cfg = cfg
            ctx.has = (bias is not None, False)
            ctx.save_for_backward(weight, bias, residual, out if act != "none" else None, *xs)
            return out
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
        xs_k = xs
        if PRECISION == "3xtf32":
            xs_k = [tf32_split(x) for x in xs]             # per pixel: hi | lo halves, each padded to 4 channels
            table, kcol = ordered_table([_pad4(c) for c in src_C], kh, kw, dev, split3=True)
            wmat, wcols = gemm_weight3(weight.detach(), src_C, w_C)
        else:
            patchable = src_C == w_C and _patch_ok(Cin, N, Hin, Win, B, kh, kw, stride, pad, reflect, len(xs), ups[0], 1)
            use_rows = (not patchable) and len(src_C) == len(w_C) and _fwd_rows_ok(src_C, ups, N, kh, kw, stride, pad, reflect, Hin, Win, Ho, Wo)
            if use_rows:
                # TMA-row operand: dense full-resolution sources of whole 32-channel blocks
                xs_k = [upsample2x(x) if u else (pad_channels(x, 32) if x.shape[1] % 32 else x) for x, u in zip(xs, ups)]
                ups = [False] * len(xs_k)
                ctx.dense_sources = xs_k          # the weight gradient reads the same materialised sources
            elif len(xs) > 1 and any(c % 4 for c in src_C):
                # a source with a ragged channel count (the 1-channel disparity of the iconv layers) is zero-padded to whole 16-byte
                # chunks: its K blocks then take the asynchronous copy path instead of eight dependent scalar loads per thread
                xs_k = [pad_channels(x, _pad4(x.shape[1])) if x.shape[1] % 4 else x for x in xs]
            src_k = [x.shape[1] for x in xs_k]
            table, kcol = ordered_table(src_k, kh, kw, dev)
            wmat, wcols = gemm_weight(weight.detach(), src_k, w_C)
        nkb = table.shape[0] // 8
        ks = _ksplit(B * Ho * Wo, N, nkb)
        has_epi = bias is not None or residual is not None or act != "none"
        if ks > 1 and has_epi and (N % 4 or not SPLIT_EPILOGUE):
            ks = 1
        # split-K layers with an epilogue: the partial tiles are summed without it and one small pass finishes in place
        finish = ks > 1 and has_epi
        out = torch.empty((B, N, Ho, Wo), dtype=torch.float32, device=dev, memory_format=CL)
        if ks > 1:
            out.zero_()
        a = _lib.ConvArgs()
        a.ksplit = ks
        a.acc_scale = _acc_scale()
        _fill_sources(a, xs_k, ups)
        a.B, a.Hin, a.Win, a.Ho, a.Wo, a.N = B, Hin, Win, Ho, Wo, N
        a.stride, a.pad, a.reflect = stride, pad, int(reflect)
        a.weight, a.w_row, a.w_cols = ptr(wmat), wmat.stride(0), wcols
        a.table, a.nkb = ptr(table), table.shape[0] // 8
        a.ntaps, a.kw = kh * kw, kw
        a.kcol, a.l1_gather = (ptr(kcol) if kcol is not None else None), int(L1_GATHER and kcol is not None)
        if residual is not None:
            residual = residual if residual.is_contiguous(memory_format=CL) else residual.contiguous(memory_format=CL)
        if not finish:
            a.bias = ptr(bias.detach()) if bias is not None else None
            a.residual = ptr(residual) if residual is not None else None
            a.act = ACT[act]
        a.out = ptr(out)
        STATS_FUSED[0] = False
        if cfg.get("bn_stats") and ks == 1 and N % 4 == 0 and N <= 2048:
            from .functional import bn_stats_pointer
            a.stats = bn_stats_pointer(N, dev)
            STATS_FUSED[0] = True
        if DBG_STAMPS is not None:
            a.dbg = ptr(DBG_STAMPS)
        a.dbg_skip = DBG_SKIP
        if src_C == w_C and wcols == kh * kw * Cin and _patch_ok(Cin, N, Hin, Win, B, kh, kw, stride, pad, reflect, len(xs), ups[0], ks):
            a.patch, a.patch_desc_mode = 1 + PATCH_TILE_ROWS, PATCH_DESC_MODE
            _patch_taps(a, TAPS_3X3, 2, (1, 1))
        elif PRECISION == "tf32" and use_rows:
            a.rows, a.rows_wv = FWD_ROWS, Wo
        tag = (B * Ho * Wo, N, table.shape[0] * 4, kh, stride, tuple(src_C), tuple(int(u) for u in cfg["ups"]), int(bool(reflect)), ks)
        check(_launch("conv_fwd", out, lambda: _lib.lib().jpb_conv2d_fwd(C.byref(a), stream_of(out)), tag), "jpb_conv2d_fwd")
        if finish:
            check(_launch("conv_bias_act", out, lambda: _lib.lib().jpb_bias_act(
                ptr(out), ptr(bias.detach()) if bias is not None else None, ptr(residual) if residual is not None else None,
                B * Ho * Wo, N, ACT[act], stream_of(out))), "jpb_bias_act")
        ctx.cfg = cfg
        ctx.has = (bias is not None, residual is not None)
        ctx.save_for_backward(weight, bias, residual, out if act != "none" else None, *xs)
        return out

    @staticmethod
    def backward(ctx, gy):
        cfg = ctx.cfg
        weight, bias, residual, out, *xs = ctx.saved_tensors
        ups, stride, pad, reflect, act = cfg["ups"], cfg["stride"], cfg["pad"], cfg["reflect"], cfg["act"]
        N = weight.shape[0]
        from .functional import direct_grad_target
        bt = direct_grad_target(bias) if bias is not None else None
        dz, gb = act_bwd(gy, out, act, bias is not None, bt)
        if bt is not None:
            gb = None
        gr = dz if residual is not None else None
        if SMALLN and _is_smalln(weight, xs, stride, pad, residual):
            gw, gx = smalln_bwd(xs[0], ups[0], _cl(dz), weight, reflect, ctx.needs_input_grad[1], ctx.needs_input_grad[4])
            return (None, gw, gb, gr, gx)
        dzp = dz
        if N % 4:                                    # e.g. the 6-channel pose head: pad dz so rows are whole 16-byte chunks
            dzp = F.pad(dz, (0, 0, 0, 0, 0, _pad4(N) - N)).contiguous(memory_format=CL)
        gxs = [None] * len(xs)
        target = direct_grad_target(weight)
        want_dx = any(ctx.needs_input_grad[4:])
        if want_dx:
            gxs = conv_dgrad(dzp, weight, xs, ups, stride, pad, reflect, ctx.needs_input_grad[4:])
        gw = conv_wgrad(dzp, weight, xs, ups, stride, pad, reflect, target=target, dense=getattr(ctx, "dense_sources", None)) if ctx.needs_input_grad[1] else None
        return (None, gw, gb, gr) + tuple(gxs)


# ---- TMA-patch kernel (csrc/conv_tc.cu: conv_tc_patch_kernel) ----
PATCH = int(_os.environ.get("JPB_CONV_PATCH", "1"))            # 0: always the gather kernels
PATCH_DESC_MODE = int(_os.environ.get("JPB_CONV_PATCH_DESC", "0"))
PATCH_MIN_TILES = int(_os.environ.get("JPB_CONV_PATCH_MIN_TILES", "96"))
PATCH_TILE_ROWS = int(_os.environ.get("JPB_CONV_PATCH_TR", "0"))   # 0: the library picks 1 or 2 tiles per CTA; 1 / 2: forced (tests)


def _patch_ok(C, N, H, W, B, kh, kw, stride, pad, reflect, nsrc, up, ks):
    """3x3 / stride 1 / zero pad 1 over one dense source, channels in whole 32-blocks, rows in whole 8-pixel tiles, enough
    16x8 tiles to occupy the machine (smaller layers keep the split-K gather kernels)."""
    if not PATCH or PRECISION != "tf32" or ks > 1:
        return False
    if not (kh == 3 and kw == 3 and stride == 1 and pad == 1 and not reflect and nsrc == 1 and not up):
        return False
    if C % 32 or N % 16 or W % 8:
        return False
    nt = 16
    while nt < N and nt < 256:
        nt *= 2
    return ((N + nt - 1) // nt) * B * ((H + 15) // 16) * (W // 8) >= PATCH_MIN_TILES


def _patch_taps(a, taps, halo, org):
    """Fill the tap geometry of a patch-mode launch: ``taps`` = [(dy, dx)] offsets inside the patch (rows, pixels)."""
    a.patch_ntaps, a.patch_halo, a.patch_org_y, a.patch_org_x = len(taps), halo, org[0], org[1]
    for t, (dy, dx) in enumerate(taps):
        a.patch_tapoff[t] = (dy * 2048 + dx * 128) >> 4


TAPS_3X3 = [(ky, kx) for ky in range(3) for kx in range(3)]
# the 7x7 / stride-2 / pad-3 stems in space-to-depth form (csrc/elementwise.cu: stem_s2d_kernel): 4 row taps (oy-2 .. oy+1) x 2
# position taps (ox-1, ox+1 in x3's shifted coordinates), patch origin (tile row - 2, tile pixel - 1)
TAPS_STEM = [(ai, bi) for ai in range(4) for bi in (0, 2)]
STEM_PATCH = int(_os.environ.get("JPB_CONV_STEM_PATCH", "1"))
_STEM_INDEX: dict = {}


def _stem_ok(weight, xs, ups, stride, pad, reflect, ks):
    N, Cin, kh, kw = weight.shape
    if not (STEM_PATCH and PATCH and PRECISION == "tf32" and kh == 7 and kw == 7 and stride == 2 and pad == 3 and not reflect and len(xs) == 1 and not ups[0]):
        return False
    B, Cp, H, W = xs[0].shape
    return Cp in (4, 8) and Cin <= Cp and H % 2 == 0 and W % 2 == 0 and (W // 2) % 8 == 0 and N % 16 == 0


def stem_weight(weight, Cp):
    """[N, 8 taps * 8*Cp] GEMM operand of the space-to-depth stem: column (t, dy, dx, c) = w[n, c, 2a+dy+3, 2b+dx+3] for tap
    t = (a, b), a in -2..1, b in {-2, 0}; zero where the 7x7 window has no such element or c is a padding channel."""
    N, Cin, kh, kw = weight.shape
    key = (Cin, Cp, str(weight.device))
    idx = _STEM_INDEX.get(key)
    if idx is None:
        cols = []
        zero = Cin * 49                     # index of an appended zero column
        for a_ in (-2, -1, 0, 1):
            for b_ in (-2, 0):
                for dy in range(2):
                    for dx in range(4):
                        ky, kx = 2 * a_ + dy + 3, 2 * b_ + dx + 3
                        for c in range(Cp):
                            cols.append((c * 49 + ky * 7 + kx) if (0 <= ky < 7 and 0 <= kx < 7 and c < Cin) else zero)
        idx = _STEM_INDEX[key] = torch.tensor(cols, dtype=torch.long, device=weight.device)
    w = torch.cat([weight.detach().reshape(N, Cin * 49), torch.zeros(N, 1, dtype=weight.dtype, device=weight.device)], 1)
    return w.index_select(1, idx).contiguous()


def stem_forward(x, weight, bias, act, bn_stats):
    """7x7 / stride-2 / pad-3 stem as an 8-tap patch convolution over the space-to-depth input (no gather)."""
    x = _cl(x)
    B, Cp, H, W = x.shape
    N = weight.shape[0]
    Ho, Wo = H // 2, W // 2
    dev = x.device
    x3 = torch.empty((B, 8 * Cp, Ho, Wo + 1), dtype=torch.float32, device=dev, memory_format=CL)
    check(_launch("stem_s2d", x, lambda: _lib.lib().jpb_stem_s2d(ptr(x), ptr(x3), B, H, W, Cp, stream_of(x))), "jpb_stem_s2d")
    wmat = stem_weight(weight, Cp)
    out = torch.empty((B, N, Ho, Wo), dtype=torch.float32, device=dev, memory_format=CL)
    a = _lib.ConvArgs()
    a.acc_scale = _acc_scale()
    _fill_sources(a, [x3], [False])
    a.B, a.Hin, a.Win, a.Ho, a.Wo, a.N = B, Ho, Wo, Ho, Wo, N
    a.stride, a.pad, a.reflect = 1, 0, 0
    a.weight, a.w_row, a.w_cols = ptr(wmat), wmat.stride(0), wmat.shape[1]
    table = chunk_table([8 * Cp], 1, 1, dev)          # unused by the patch kernel (argument check only)
    a.table, a.nkb = ptr(table), table.shape[0] // 8
    a.ntaps, a.kw = len(TAPS_STEM), 2
    a.bias = ptr(bias.detach()) if bias is not None else None
    a.act = ACT[act]
    a.out = ptr(out)
    STATS_FUSED[0] = False
    if bn_stats and N % 4 == 0:
        from .functional import bn_stats_pointer
        a.stats = bn_stats_pointer(N, dev)
        STATS_FUSED[0] = True
    a.patch, a.patch_desc_mode = 1 + PATCH_TILE_ROWS, PATCH_DESC_MODE
    _patch_taps(a, TAPS_STEM, 3, (2, 1))
    a.dbg_skip = DBG_SKIP
    tag = (B * Ho * Wo, N, weight.shape[1] * 49, 7, 2, (Cp,), (0,), 0, 1)
    check(_launch("conv_fwd", out, lambda: _lib.lib().jpb_conv2d_fwd(C.byref(a), stream_of(out)), tag), "jpb_conv2d_fwd(stem)")
    return out


STATS_FUSED = [False]   # set by the last forward launch: its epilogue accumulated the BatchNorm statistics of its output


def conv2d_tc(xs, ups, weight, bias, stride, pad, reflect, act, residual, bn_stats=False):
    cfg = dict(ups=tuple(bool(u) for u in ups), stride=stride, pad=pad, reflect=bool(reflect), act=act, bn_stats=bool(bn_stats))
    return _ConvTC.apply(cfg, weight, bias, residual, *xs)
