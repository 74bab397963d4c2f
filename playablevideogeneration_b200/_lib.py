"""ctypes binding of the C ABI declared in include/pvg_b200.h (the drop-in boundary)."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int32, c_int64, c_void_p, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PVG_LIB") or os.path.join(_HERE, "libpvg_b200.so")      # PVG_LIB: an alternative build (A/B timing)

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_LSTM = 0, 1, 2, 3, 4, 5
ALGO_AUTO, ALGO_SIMT, ALGO_UMMA, ALGO_UMMA_PERSISTENT = 0, 1, 2, 3
CORR_BF16, CORR_FP16, CORR_FP16_ALL = 0, 1, 2


class ConvDesc(Structure):
    """struct pvg_conv_desc (include/pvg_b200.h)."""
    _fields_ = [("N", c_int32), ("H", c_int32), ("W", c_int32), ("Cin", c_int32), ("Cout", c_int32),
                ("R", c_int32), ("S", c_int32), ("pad", c_int32), ("act", c_int32), ("slope", c_float),
                ("algo", c_int32), ("nprod", c_int32), ("corr_fmt", c_int32)]


class ConcatDesc(Structure):
    """struct pvg_concat_desc (include/pvg_b200.h)."""
    MAX_PARTS = 6
    _fields_ = [("N", c_int32), ("H", c_int32), ("W", c_int32), ("Cpad", c_int32), ("nparts", c_int32),
                ("c", c_int32 * 6), ("is_vec", c_int32 * 6), ("bstride", c_int64 * 6), ("src", c_void_p * 6)]


P = c_void_p
# name -> (argtypes); every function returns int (0 = ok) except pvg_last_error / pvg_version / pvg_has_umma
_SIGNATURES = {
    "pvg_conv2d_fwd": [POINTER(ConvDesc), P, P, P, P, P, P, P],
    "pvg_conv2d_stem_planes": [POINTER(ConvDesc), P, P, P, P, P, P],
    "pvg_conv2d_fwd_planes": [POINTER(ConvDesc), P, P, P, P, P, P, P, c_int, P, P],
    "pvg_conv2d_wgrad_planes": [POINTER(ConvDesc), c_int, P, P, P, P, P, c_int, P],
    "pvg_conv2d_wgrad_small": [POINTER(ConvDesc), c_int, P, P, P, P, c_int, P],
    "pvg_unpack_dw": [P, c_int, c_int, c_int, c_int, c_int, P, c_int, P],
    "pvg_amax": [P, c_int64, P, P],
    "pvg_split_16_scaled": [P, P, c_int64, P, P, P],
    "pvg_act_bwd_split_16_scaled": [P, P, c_int, c_float, P, P, c_int64, P, P, P],
    "pvg_pack_conv_weight": [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P],
    "pvg_conv2d_wgrad": [POINTER(ConvDesc), c_int, P, P, P, P],
    "pvg_conv2d_wgrad_umma": [POINTER(ConvDesc), c_int, P, P, P, P, P, P, c_int, P],
    "pvg_channel_sum": [P, c_int64, c_int, P, P, P],
    "pvg_split_tf32": [P, P, P, c_int64, P],
    "pvg_split_16": [P, P, c_int64, c_int, P],
    "pvg_pack_16x2": [P, P, P, c_int64, c_int, P],
    "pvg_act_bwd_split_16": [P, P, c_int, c_float, P, P, c_int64, c_int, P],
    "pvg_act_bwd": [P, P, c_int, c_float, P, c_int64, P],
    "pvg_act_bwd_tap": [P, P, c_int, c_float, P, c_int, c_int64, P, P, P],
    "pvg_act_bwd_tap_split_16_scaled": [P, P, c_int, c_float, P, P, c_int, c_int64, P, P, P, P, P],
    "pvg_act_bwd_split": [P, P, c_int, c_float, P, P, P, c_int64, P],
    "pvg_bn_stats": [P, c_int, c_int, c_int, c_int, P, P],
    "pvg_pool2_stats": [P, c_int, c_int, c_int, c_int, P, c_int, P, P],
    "pvg_bn_finalize": [P, c_int64, c_int, c_int, c_float, c_float, P, P, P, P, P],
    "pvg_bn_eval_prepare": [P, P, c_int, c_float, P, P, P],
    "pvg_bn_apply": [P, c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, c_float, P, P],
    "pvg_bn_apply_ex": [P, c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, c_float, P, P, c_int, P, c_int, c_int, P],
    "pvg_bn_finalize_apply_ex": [P, c_int, c_int, c_int, c_int, P, c_int64, c_float, c_float, P, P, P, P, P, P, P, c_int, c_float,
                                 P, P, c_int, P, c_int, c_int, P],
    "pvg_maxpool2_fwd_ex": [P, c_int, c_int, c_int, c_int, P, P, c_int, P],
    "pvg_resize_bilinear_ex": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, P, c_int, P, c_int, P],
    "pvg_concat_pad": [POINTER(ConcatDesc), P, P, c_int, P, c_int, P],
    "pvg_bn_bwd_reduce": [P, P, P, c_int, c_int, c_int, c_int, P, P, c_int, c_float, P, P],
    "pvg_bn_bwd_apply": [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, c_float, P, c_int, c_int, P, P, P, P, P],
    "pvg_bn_bwd_apply_ex": [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, c_float, P, c_int, c_int, P, P, P, P, c_int, P, P],
    "pvg_pack_conv_weight_ex": [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P],
    "pvg_bn_finalize_apply": [P, c_int, c_int, c_int, c_int, P, c_int64, c_float, c_float, P, P, P, P, P, P, P, c_int, c_float,
                              P, P],
    "pvg_bn_bwd_params": [P, c_int, c_int, P, P, P],
    "pvg_upsample2x_fwd": [P, c_int, c_int, c_int, c_int, P, P],
    "pvg_upsample2x_bwd": [P, c_int, c_int, c_int, c_int, P, P],
    "pvg_resize_bilinear": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, P],
    "pvg_maxpool2_fwd": [P, c_int, c_int, c_int, c_int, P, P],
    "pvg_maxpool2_bwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "pvg_lstm_fwd": [P, P, c_int64, c_int, P, P, P],
    "pvg_convlstm_step": [POINTER(ConvDesc), P, P, P, P, P, P, P, P],
    "pvg_lstm_bwd_act": [P, P, P, P, P, c_int64, c_int, P, P, P],
    "pvg_lstm_bwd": [P, P, P, P, P, c_int64, c_int, P, P, P],
    "pvg_absdiff_mean_fwd": [P, P, c_int, c_int64, P, P],
    "pvg_absdiff_mean_bwd": [P, P, P, c_int, c_int64, P, P],
    "pvg_resample_u8": [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P],
    "pvg_frames_u8_to_nhwc": [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, P, P],
    "pvg_sqdiff_mean": [P, P, c_int, c_int, c_int, c_int64, c_int, c_int, P, P],
    "pvg_frames_to_u8": [P, c_int64, P, P, P],
    "pvg_adam_step": [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, c_float, P],
    "pvg_adam_step_dev": [P, P, P, P, c_int64, P, P],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["pvg_last_error", "pvg_version", "pvg_has_umma"])

_lib = None
launch_count = 0          # number of C-ABI calls that launched kernels (bench.py reports it as gpu_launches)


class PvgError(RuntimeError):
    pass


def load():
    """Loads libpvg_b200.so.  Raises (never falls back) when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise PvgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU / eager fallback for the CADDY hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    lib.pvg_last_error.restype = ctypes.c_char_p
    lib.pvg_last_error.argtypes = []
    lib.pvg_version.restype = c_int
    lib.pvg_has_umma.restype = c_int
    _lib = lib
    return lib


def call(name: str, *args):
    """Invokes one C-ABI entry point and raises PvgError with pvg_last_error() on a non-zero return."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PvgError(f"{name} failed ({rc}): {lib.pvg_last_error().decode()}")
    launch_count += 1
    return rc
