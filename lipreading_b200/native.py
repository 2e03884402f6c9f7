"""ctypes binding of liblr_b200.so (the C-ABI in include/lr_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblr_b200.so")

LR_RNN_TANH, LR_RNN_GRU, LR_RNN_LSTM = 0, 1, 2
RNN_MODES = {"RNN": LR_RNN_TANH, "GRU": LR_RNN_GRU, "LSTM": LR_RNN_LSTM}
RNN_GATES = {"RNN": 1, "GRU": 3, "LSTM": 4}


class NativeError(RuntimeError):
    pass


_lib = None

_vp, _i, _sz, _i64, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int64, ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/lr_b200.h one to one.
SIGNATURES = {
    "lr_abi_version": (_i, []),
    "lr_last_error": (ctypes.c_char_p, []),
    "lr_launch_count": (_u64, []),
    "lr_ctc_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "lr_ctc_fwd_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "lr_ctc_greedy_decode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "lr_scale_rows": (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    "lr_proj_logsoftmax_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "lr_proj_tc5_supported": (_i, [_i, _i, _i]),
    "lr_logsoftmax_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "lr_proj_logsoftmax_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "lr_rnn_saved_per_unit": (_i, [_i]),
    "lr_rnn_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "lr_rnn_fwd": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lr_rnn_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lr_rnn_cluster_supported": (_i, [_i, _i]),
    "lr_rnn_cluster_fwd": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "lr_rnn_cluster_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lr_rnn_grid_supported": (_i, [_i, _i, _i]),
    "lr_rnn_grid_workspace": (_sz, [_i, _i, _i]),
    "lr_rnn_grid_fwd": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lr_rnn_grid_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lr_collate_pad_f64": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "lr_rect_geometry": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "lr_warp256": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "lr_posmap_gather": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "lr_mouth_crop": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lr_attn_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lr_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lr_attn_scores_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lr_attn_scores_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lr_tapgemm": (_i, [_vp, _vp]),                       # (const lr_tapgemm_desc*, stream): prnet_tc5._Desc
    "lr_pack_image16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
}


class TapGemmDesc(ctypes.Structure):
    """lr_tapgemm_desc of include/lr_b200.h, field for field"""
    _fields_ = [("a", ctypes.c_void_p), ("rows", ctypes.c_longlong), ("C", ctypes.c_int),
                ("Hp", ctypes.c_int), ("Wp", ctypes.c_int), ("vy0", ctypes.c_int), ("vx0", ctypes.c_int),
                ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("w", ctypes.c_void_p), ("w_pitch", ctypes.c_int), ("Kg", ctypes.c_int), ("Kt", ctypes.c_int),
                ("n_phases", ctypes.c_int), ("n_groups", ctypes.c_int), ("tap_off", ctypes.POINTER(ctypes.c_int32)),
                ("Cout_pad", ctypes.c_int), ("Cout", ctypes.c_int),
                ("alpha", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("gamma", ctypes.c_void_p),
                ("act", ctypes.c_int), ("mode", ctypes.c_int),
                ("out", ctypes.c_void_p), ("oHp", ctypes.c_int), ("oWp", ctypes.c_int), ("oC", ctypes.c_int),
                ("opy", ctypes.c_int), ("opx", ctypes.c_int),
                ("res", ctypes.c_void_p), ("resC", ctypes.c_int),
                ("aux", ctypes.c_void_p), ("aHp", ctypes.c_int), ("aWp", ctypes.c_int), ("aC", ctypes.c_int),
                ("apad", ctypes.c_int), ("out_scale", ctypes.c_float), ("flags", ctypes.c_int), ("pack", ctypes.c_int)]


# include/lr_b200_diag.h — measurement hooks and micro-benchmarks, only in liblr_b200_diag.so (tools/ use them)
DIAG_LIB_PATH = os.path.join(_HERE, "liblr_b200_diag.so")
DIAG_SIGNATURES = {
    "lr_conv3d_set_debug": (None, [_vp]),
    "lr_conv3d_set_debug_skip": (None, [_i]),
    "lr_umma_microbench": (ctypes.c_longlong, [_i] * 10 + [_vp]),
    "lr_umma_pattern_bench": (ctypes.c_longlong, [_vp, _vp, _vp, _i, _i, _vp]),
    "lr_umma_issue_bench": (ctypes.c_longlong, [_i] * 8 + [_vp]),
}


def use_diag_lib():
    """tools/ only: make lib() load the diagnostics build (same entry points + the hooks of lr_b200_diag.h)."""
    global LIB_PATH, _lib
    assert _lib is None, "call use_diag_lib() before the first native.lib()"
    LIB_PATH = DIAG_LIB_PATH
    SIGNATURES.update(DIAG_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle.  Raises NativeError when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "liblr_b200.so not built (%s). Run `python -m lipreading_b200.build` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)       # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def register(name, restype, argtypes):
    SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def check(rc, what):
    if rc != 0:
        raise NativeError("%s failed (%d): %s" % (what, rc, lib().lr_last_error().decode()))


def launch_count():
    return int(lib().lr_launch_count())


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeError("lipreading_b200 ops run on CUDA tensors only (got %s); there is no CPU path" % t.device)


def cont(t, dtype=None):
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()
