"""ctypes binding of ``libjpb200.so`` (the C ABI declared in ``include/jpb200.h``).

The product path has exactly one backend: the sm_100a CUDA library built in-tree by
``jperceiver_b200.build``.  If it is missing, :func:`lib` raises — there is no CPU or PyTorch
fallback.  (``use_library`` exists so the test-suite can install the host *emulation* build of the
same sources, see ``tests/emu``; nothing in the package calls it.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libjpb200.so")
MAX_SRC = 4

_handle = None
_emulated = False
launches = 0  # number of kernel launches issued through this binding (bench.py reads it)


class JpbError(RuntimeError):
    pass


class PhotoArgs(C.Structure):
    _fields_ = [
        ("target", C.c_void_p), ("src", C.c_void_p * MAX_SRC), ("T", C.c_void_p * MAX_SRC),
        ("noise", C.c_void_p * MAX_SRC), ("disp", C.c_void_p), ("K", C.c_void_p), ("invK", C.c_void_p),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("hs", C.c_int), ("ws", C.c_int), ("F", C.c_int),
        ("automask", C.c_int), ("min_disp", C.c_float), ("max_disp", C.c_float), ("noise_scale", C.c_float),
        ("seed", C.c_uint64), ("stream", C.c_uint64), ("step", C.c_void_p), ("loss_sum", C.c_void_p), ("min_index", C.c_void_p),
        ("winner", C.c_void_p), ("warped", C.c_void_p * MAX_SRC), ("ident_err", C.c_void_p), ("ident_mode", C.c_int),
    ]


class PhotoGrad(C.Structure):
    _fields_ = [
        ("grad_out", C.c_void_p), ("inv_count", C.c_float), ("winner", C.c_void_p), ("grad_disp", C.c_void_p),
        ("grad_T", C.c_void_p * MAX_SRC),
    ]


class Pyramid(C.Structure):
    _fields_ = [("nlev", C.c_int), ("level", C.c_void_p * 4)]


class ScaleLabelArgs(C.Structure):
    _fields_ = [
        ("label", C.c_void_p), ("K3", C.c_void_p), ("k_stride", C.c_int), ("k_row", C.c_int), ("Tr", C.c_void_p),
        ("quad", C.c_void_p), ("out", C.c_void_p), ("B", C.c_int), ("occ", C.c_int), ("Hf", C.c_int), ("Wf", C.c_int),
        ("mode", C.c_int), ("align_corners", C.c_int), ("z_offset", C.c_float), ("cam_height", C.c_float),
    ]


class ScaleLossArgs(C.Structure):
    _fields_ = [
        ("disp", C.c_void_p), ("label", C.c_void_p), ("B", C.c_int), ("hs", C.c_int), ("ws", C.c_int), ("Hf", C.c_int),
        ("Wf", C.c_int), ("crop", C.c_int), ("min_disp", C.c_float), ("max_disp", C.c_float), ("acc", C.c_void_p),
        ("weight", C.c_float), ("grad_out", C.c_void_p), ("grad_disp", C.c_void_p),
    ]


class BevArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("stride_b", C.c_longlong), ("stride_c", C.c_longlong), ("stride_p", C.c_longlong),
        ("label", C.c_void_p), ("sdf", C.c_void_p), ("B", C.c_int), ("occ", C.c_int), ("w_fg", C.c_float),
        ("loss_weight", C.c_float), ("loss2_weight", C.c_float), ("acc", C.c_void_p), ("region", C.c_int), ("loss_sum", C.c_int),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * 3), ("src_C", C.c_int * 3), ("src_H", C.c_int * 3), ("src_W", C.c_int * 3), ("src_up", C.c_int * 3),
        ("nsrc", C.c_int), ("B", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int), ("N", C.c_int),
        ("stride", C.c_int), ("pad", C.c_int), ("reflect", C.c_int), ("weight", C.c_void_p), ("w_row", C.c_longlong),
        ("w_cols", C.c_int), ("table", C.c_void_p), ("nkb", C.c_int), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("act", C.c_int), ("out", C.c_void_p), ("in_div", C.c_int), ("nt", C.c_int), ("scatter", C.c_int),
        ("dst", C.c_void_p * 3), ("dst_C", C.c_int * 3), ("dst_H", C.c_int * 3), ("dst_W", C.c_int * 3), ("dst_up", C.c_int * 3),
        ("ndst", C.c_int), ("fold_pad", C.c_int), ("fold_reflect", C.c_int), ("fold_H", C.c_int), ("fold_W", C.c_int),
        ("ntaps", C.c_int), ("kw", C.c_int), ("ksplit", C.c_int), ("kcol", C.c_void_p), ("l1_gather", C.c_int),
        ("dbg", C.c_void_p), ("dbg_skip", C.c_int), ("stats", C.c_void_p), ("acc_scale", C.c_float), ("patch", C.c_int), ("patch_ntaps", C.c_int), ("patch_halo", C.c_int), ("patch_org_y", C.c_int),
        ("patch_org_x", C.c_int), ("patch_tapoff", C.c_int * 16), ("patch_desc_mode", C.c_int), ("rows", C.c_int), ("rows_wv", C.c_int),
        ("dst_mul", C.c_int), ("dst_oy", C.c_int), ("dst_ox", C.c_int),
    ]


class ConvWgradArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * 3), ("src_C", C.c_int * 3), ("src_H", C.c_int * 3), ("src_W", C.c_int * 3), ("src_up", C.c_int * 3),
        ("nsrc", C.c_int), ("B", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int), ("N", C.c_int),
        ("stride", C.c_int), ("pad", C.c_int), ("reflect", C.c_int), ("table", C.c_void_p), ("nchunks", C.c_int),
        ("dy", C.c_void_p), ("dw", C.c_void_p), ("w_row", C.c_longlong), ("w_cols", C.c_int), ("splits", C.c_int),
        ("dbg", C.c_void_p), ("accumulate", C.c_int), ("acc_scale", C.c_float), ("dy_pitch", C.c_int),
        ("rows", C.c_int), ("gflags", C.c_void_p), ("dz3", C.c_int), ("chunk_col", C.c_void_p),
    ]


class WeightT(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("N", C.c_int), ("Cin", C.c_int), ("taps", C.c_int), ("block_start", C.c_int)]


class DepthEvalArgs(C.Structure):
    _fields_ = [("disp", C.c_void_p), ("gt", C.c_void_p), ("B", C.c_int), ("h", C.c_int), ("w", C.c_int), ("gh", C.c_int),
                ("gw", C.c_int), ("min_disp", C.c_float), ("max_disp", C.c_float), ("min_depth", C.c_float), ("max_depth", C.c_float),
                ("crop", C.c_int * 4), ("fixed_scale", C.c_float), ("work", C.c_void_p), ("count", C.c_void_p), ("out", C.c_void_p)]


class ResizeArgs(C.Structure):
    _fields_ = [("src", C.c_void_p), ("tmp", C.c_void_p), ("dst", C.c_void_p), ("dst_f", C.c_void_p), ("B", C.c_int), ("Hin", C.c_int),
                ("Win", C.c_int), ("Hout", C.c_int), ("Wout", C.c_int), ("kx", C.c_void_p), ("bx", C.c_void_p), ("ksx", C.c_int),
                ("ky", C.c_void_p), ("by", C.c_void_p), ("ksy", C.c_int), ("flip", C.c_void_p)]


class JitterArgs(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst_f", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("order", C.c_void_p), ("factor", C.c_void_p), ("hue_shift", C.c_void_p), ("enable", C.c_void_p), ("lsum", C.c_void_p)]


class AdamArgs(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("grad_scale", C.c_float), ("max_norm", C.c_float), ("normsq", C.c_void_p), ("step", C.c_void_p)]


EMU_MISSING = ("jpb_conv2d_fwd", "jpb_conv2d_wgrad", "jpb_conv_set_pair")   # tcgen05/TMA entry points do not exist in the host-emulation build


def _declare(h):
    h.jpb_abi_version.restype = C.c_int
    h.jpb_build_info.restype = C.c_char_p
    h.jpb_bn_workspace_doubles.restype = C.c_longlong
    h.jpb_bn_workspace_doubles.argtypes = [C.c_int]
    h.jpb_bn_stats_accumulator.restype = C.c_void_p
    h.jpb_bn_stats_accumulator.argtypes = [C.c_void_p]
    for name in dir(_Signatures):
        if name.startswith("jpb_"):
            if not hasattr(h, name) and name in EMU_MISSING:
                continue
            fn = getattr(h, name)
            fn.restype = C.c_int
            fn.argtypes = getattr(_Signatures, name)


class _Signatures:
    """argtypes of every int-returning entry point (kept next to include/jpb200.h)."""
    P, I, F, V = C.c_void_p, C.c_int, C.c_float, C.c_void_p
    jpb_photometric_fwd = [C.POINTER(PhotoArgs), V]
    jpb_photometric_bwd = [C.POINTER(PhotoArgs), C.POINTER(PhotoGrad), V]
    jpb_photometric_set_variant = [I]
    jpb_photometric_get_variant = []
    jpb_photometric_set_bwd_variant = [I]
    jpb_finalize = [P, P, F, P, I, V]
    jpb_weight_flipT = [P, I, I, V]
    jpb_cct_select_fwd = [P] * 10 + [I, I, I, I, V]
    jpb_cct_select_bwd = [P] * 14 + [I, I, I, I, V]
    jpb_cct_combine_fwd = [P] * 6 + [I, I, I, I, V]
    jpb_cct_combine_bwd = [P] * 9 + [I, I, I, I, V]
    jpb_cvp_mlp_fwd = [P] * 7 + [I, I, I, V]
    jpb_cvp_mlp_bwd = [P] * 13 + [I, I, I, V]
    jpb_image_prep = [P, P, P, I, I, I, I, I, I, V]
    jpb_dropout = [P, P, P, C.c_longlong, F, C.c_uint64, C.c_uint64, P, V]
    jpb_pose_head_fwd = [P, P, P, I, I, I, I, V]
    jpb_pose_head_bwd = [P, P, P, I, I, I, I, V]
    jpb_area_pyramid = [P, I, I, I, C.POINTER(Pyramid), V]
    jpb_smooth_fwd = [P, P, I, I, I, I, F, P, P, V]
    jpb_smooth_bwd = [P, P, I, I, I, I, F, P, P, P, V]
    jpb_scale_label = [C.POINTER(ScaleLabelArgs), V]
    jpb_scale_loss_fwd = [C.POINTER(ScaleLossArgs), V]
    jpb_scale_loss_bwd = [C.POINTER(ScaleLossArgs), V]
    jpb_signed_distance = [P, I, I, P, P, V]
    jpb_bev_loss_fwd = [C.POINTER(BevArgs), P, V]
    jpb_bev_loss_bwd = [C.POINTER(BevArgs), P, P, V]
    jpb_l1_mean_fwd = [P, P, C.c_longlong, P, V]
    jpb_l1_mean_bwd = [P, P, C.c_longlong, P, P, P, V]
    jpb_sumsq = [P, C.c_longlong, P, V]
    jpb_conv2d_fwd = [C.POINTER(ConvArgs), V]
    jpb_conv2d_wgrad = [C.POINTER(ConvWgradArgs), V]
    jpb_conv_set_pair = [I]
    jpb_act_bwd = [P, P, P, C.c_longlong, I, I, P, V]
    jpb_bias_act = [P, P, P, C.c_longlong, I, I, V]
    jpb_tf32_split = [P, P, C.c_longlong, I, V]
    jpb_stem_s2d = [P, P, I, I, I, I, V]
    jpb_upsample2x = [P, P, I, I, I, I, V]
    jpb_pad_channels = [P, P, C.c_longlong, I, I, V]
    jpb_sum_n = [P, I, P, C.c_longlong, V]
    jpb_conv3x3_smalln_fwd = [P, P, P, P, P, I, I, I, I, I, I, I, I, V]
    jpb_conv3x3_smalln_bwd = [P, P, P, P, P, P, I, I, I, I, I, I, I, V]
    jpb_maxpool_fwd = [P, P, P, I, I, I, I, I, I, I, V]
    jpb_bn_train_fwd = [P, P, P, P, P, P, P, I, F, F, I, P, P, P, C.c_longlong, I, I, V]
    jpb_bn_eval_fwd = [P, P, P, P, P, I, P, C.c_longlong, I, V]
    jpb_bn_train_bwd = [P, P, P, P, P, I, P, P, P, P, I, P, C.c_longlong, I, V, P]
    jpb_maxpool_bwd = [P, P, P, I, I, I, I, I, I, I, V]
    jpb_maxpool_set_bwd_variant = [I]
    jpb_adam_step = [P, P, P, P, C.c_longlong, C.POINTER(AdamArgs), V]
    jpb_depth_eval = [C.POINTER(DepthEvalArgs), V]
    jpb_resize_lanczos_u8 = [C.POINTER(ResizeArgs), V]
    jpb_color_jitter_u8 = [C.POINTER(JitterArgs), V]
    jpb_bev_label_u8 = [P, P, I, I, I, I, P, P, P, V]
    jpb_bev_confusion = [P, C.c_longlong, C.c_longlong, C.c_longlong, P, I, I, P, V]


def exported_symbols():
    return ["jpb_abi_version", "jpb_build_info", "jpb_bn_workspace_doubles", "jpb_bn_stats_accumulator"] + [n for n in dir(_Signatures) if n.startswith("jpb_")]


def lib():
    """The loaded library; raises if the CUDA extension has not been built."""
    global _handle
    if _handle is None:
        if not os.path.exists(LIB_PATH):
            raise JpbError("libjpb200.so is missing (%s): build it with `python -m jperceiver_b200.build`; "
                           "jperceiver_b200 has no CPU fallback" % LIB_PATH)
        _handle = C.CDLL(LIB_PATH)
        _declare(_handle)
        v = os.environ.get("JPB_PHOTO_FWD")      # opt-in forward schedule of the photometric kernel (2 default, 3 packed)
        if v:
            check(_handle.jpb_photometric_set_variant(int(v)), "jpb_photometric_set_variant(JPB_PHOTO_FWD=%s)" % v)
        v = os.environ.get("JPB_POOL_BWD")       # backward schedule of the max-pools (measurements): 0 default, 1 scatter, 2 5x5 gather
        if v:
            check(_handle.jpb_maxpool_set_bwd_variant(int(v)), "jpb_maxpool_set_bwd_variant(JPB_POOL_BWD=%s)" % v)
    return _handle


def photo_fwd_variant():
    """Forward schedule of the photometric kernel in force (``jpb_photometric_set_variant``)."""
    return int(lib().jpb_photometric_get_variant())


def use_library(path, emulated=False):
    """TESTS ONLY: bind a different build of the same ABI (the host emulation)."""
    global _handle, _emulated
    _handle = C.CDLL(path)
    _declare(_handle)
    _emulated = emulated
    return _handle


def is_emulated():
    return _emulated


def check(status, what):
    global launches
    launches += 1
    if status != 0:
        raise JpbError("%s failed with status %d" % (what, status))


def ptr(t):
    """Device pointer of a tensor (NULL for None).  Tensors must be CUDA (or CPU under emulation)."""
    if t is None:
        return None
    if not t.is_cuda and not _emulated:
        raise JpbError("jperceiver_b200 kernels take CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)
    return t.data_ptr()


def stream_of(t):
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return None
