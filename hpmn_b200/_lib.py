"""ctypes binding of libhpmn_b200.so (include/hpmn_b200.h).

The library is the product: there is no Python / CPU fallback.  If the shared object is missing this
module raises at import time with the command that builds it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhpmn_b200.so")

HPMN_ABI_VERSION = 1
HPMN_MAX_LAYERS = 16
HPMN_MAX_HOPS = 8
HPMN_OK, HPMN_EINVAL, HPMN_EARCH, HPMN_ECUDA, HPMN_ENOMEM = 0, -1, -2, -3, -4
S_LOGLOSS, S_COVREG, S_LOSS, S_IDERR = 0, 1, 2, 3
K_FAMILIES = ["gather_fwd", "inproj_gemm", "rec_fwd", "attn_fwd", "head_fwd", "head_bwd", "attn_bwd", "rec_bwd",
              "dx_gemm", "gru_wgrad", "scatter_add", "misc"]


class hpmn_shape(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("F", C.c_int32), ("E", C.c_int32), ("H", C.c_int32),
                ("L", C.c_int32), ("hops", C.c_int32), ("front_pad", C.c_int32), ("mask_id0", C.c_int32),
                ("last_offset", C.c_int32), ("periods", C.c_int32 * HPMN_MAX_LAYERS), ("V", C.c_int64)]


class hpmn_hyper(C.Structure):
    _fields_ = [("memory_reg", C.c_float), ("l2_reg", C.c_float), ("keep_prob", C.c_float),
                ("dropout_seed", C.c_uint64), ("loss_batch", C.c_int32)]


class hpmn_outputs(C.Structure):
    _fields_ = [("scalars", C.c_void_p), ("pred", C.c_void_p), ("logit", C.c_void_p), ("w_hop0", C.c_void_p),
                ("memory", C.c_void_p)]


class HpmnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libhpmn_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/hpmn_b200.h declares: name -> (restype, argtypes)
_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
_SH, _HY, _OUT = C.POINTER(hpmn_shape), C.POINTER(hpmn_hyper), C.POINTER(hpmn_outputs)
SYMBOLS = {
    "hpmn_abi_version": (_I, []),
    "hpmn_create": (_I, [C.POINTER(_P), _I]),
    "hpmn_destroy": (None, [_P]),
    "hpmn_last_error": (C.c_char_p, [_P]),
    "hpmn_launch_count": (_L, [_P]),
    "hpmn_set_comm_stream": (_I, [_P, _P]),
    "hpmn_param_tensors": (_I, [_SH]),
    "hpmn_param_count": (_L, [_SH]),
    "hpmn_param_offsets": (_I, [_SH, C.POINTER(_L), C.POINTER(_L), _I]),
    "hpmn_workspace_bytes": (C.c_size_t, [_SH, _I]),
    "hpmn_gather_fwd": (_I, [_P, _SH, _P, _P, _P, _P]),
    "hpmn_gather_bwd": (_I, [_P, _SH, _P, _P, _P, _P, _P]),
    "hpmn_gather_bwd_multi": (_I, [_P, _SH, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, _P]),
    "hpmn_memory_fwd": (_I, [_P, _SH, _P, _P, _P, _P, _P]),
    "hpmn_memory_bwd": (_I, [_P, _SH, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_attn_fwd": (_I, [_P, _SH, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_attn_bwd": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_head_fwd": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_head_bwd": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_head_wide_param_count": (_L, [_I]),
    "hpmn_head_wide_param_offsets": (_I, [_I, C.POINTER(_L), C.POINTER(_L)]),
    "hpmn_head_wide_workspace_bytes": (C.c_size_t, [_I, _I]),
    "hpmn_head_wide_fwd": (_I, [_P, _I, _I, _HY, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_head_wide_bwd": (_I, [_P, _I, _I, _HY, _P, _P, _P, _P, _P, _P, _P]),
    "hpmn_forward": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _OUT, _P, _P]),
    "hpmn_forward_backward": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _P, _P, _I, _OUT, _P, _P]),
    "hpmn_step_host": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _P, _P, _I, _I, _OUT, _P, _P]),
    "hpmn_step_host_begin": (_I, [_P, _SH, _HY, _P, _P, _P, _P, _P, _P, _I, _I, _OUT, _P, _P]),
    "hpmn_step_host_end": (_I, [_P, _SH, _OUT, _P]),
    "hpmn_output_block": (_I, [_SH, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hpmn_prefetch_host": (_I, [_P, _SH, _P, _P, _P]),
    "hpmn_clip_adam": (_I, [_P, _P, _P, _P, _P, _L, _L, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P]),
    "hpmn_debug_wgrad": (_I, [_P, _SH, _I, _P, _L, _P, _P, _P, _I, _P]),
    "hpmn_nvls_allreduce": (_I, [_P, _P, _L, _I, _I, _I, _P]),
    "hpmn_table_grad_sources": (_I, [_P, _SH, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hpmn_debug_tcr_stamps": (_I, [C.POINTER(C.c_longlong), _I]),
    "hpmn_profile_enable": (_I, [_P, _I]),
    "hpmn_profile_read": (_I, [_P, C.POINTER(C.c_float), C.POINTER(_L)]),
    "hpmn_kernel_family_name": (C.c_char_p, [_I]),
}


def load():
    if not os.path.exists(LIB_PATH):
        # building the product is not a fallback: compile it in-tree when the toolchain is here, fail loudly otherwise
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise ImportError("%s is missing and could not be built (%s): run `python -m hpmn_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback." % (LIB_PATH, e))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.hpmn_abi_version() != HPMN_ABI_VERSION:
        raise ImportError("libhpmn_b200.so ABI %d != binding ABI %d; rebuild" % (lib.hpmn_abi_version(), HPMN_ABI_VERSION))
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc, ctx=None):
    if rc != HPMN_OK:
        msg = lib().hpmn_last_error(ctx)
        raise HpmnError(rc, msg.decode() if msg else "")
