"""ctypes binding of libstr2str_b200.so (the C ABI declared in include/str2str_b200.h).

There is no CPU fallback: if the shared library is missing, or an entry point reports an error, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstr2str_b200.so")

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); mirrors include/str2str_b200.h one to one
SIGNATURES = {
    "s2s_abi_version": (_i, []),
    "s2s_last_error": (C.c_char_p, []),
    "s2s_create": (_vp, [_vp, _vp, _vp, _vp]),
    "s2s_destroy": (None, [_vp]),
    "s2s_set_param": (_i, [_vp, C.c_char_p, _vp, _i64]),
    "s2s_finalize": (_i, [_vp, _vp]),
    "s2s_set_option": (_i, [_vp, C.c_char_p, _i]),
    "s2s_reserve": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "s2s_net_forward": (_i, [_vp, _i, _i] + [_vp] * 10),
    "s2s_trunk": (_i, [_vp, _i, _i] + [_vp] * 9),
    "s2s_embed": (_i, [_vp, _i, _i] + [_vp] * 8),
    "s2s_ipa": (_i, [_vp, _i, _i, _i] + [_vp] * 7),
    "s2s_edge_transition": (_i, [_vp, _i, _i, _i] + [_vp] * 5),
    "s2s_node_transition": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "s2s_torsion_head": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "s2s_backbone_update": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "s2s_se3_step": (_i, [_i, _i] + [_vp] * 8 + [_f, _i, _i] + [_vp] * 4),
    "s2s_se3_perturb": (_i, [_i, _i] + [_vp] * 11),
    "s2s_rng_fill": (_i, [_vp, _i, _i64, C.c_uint64, _i64, _i, _i, _vp]),
    "s2s_rng_fill_rows": (_i, [_vp, _i, _i64, C.c_uint64, _vp, _vp, _i, _vp]),
    "s2s_backbone_atoms": (_i, [_vp, _i] + [_vp] * 6),
    "s2s_linear_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "s2s_linear_tc": (_i, [_vp] * 7 + [_i] * 5 + [_vp]),
    "s2s_launch_count": (_i64, []),
    "s2s_profile_enable": (None, [_i]),
    "s2s_profile_reset": (None, []),
    "s2s_profile_list": (_i, [C.c_char_p, _i]),
    "s2s_profile_read": (_i, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(_i64)]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the native library (once). Raises if it has not been built — there is no other code path."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m str2str_b200.build` "
                    "(the CUDA library is the only implementation; there is no CPU fallback)"
                )
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the header and the binary disagree
                fn.restype = res
                fn.argtypes = args
            if lib.s2s_abi_version() != 1:
                raise RuntimeError("libstr2str_b200.so: ABI version mismatch")
            _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("str2str_b200: " + load().s2s_last_error().decode())


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor of the dtype the kernel reads (None -> NULL).  The kernels reinterpret the
    bytes, so a tensor of another dtype is an error here rather than garbage there."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("str2str_b200 kernels take CUDA tensors; got a CPU tensor (no CPU fallback exists)")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"kernel argument must be {dtype}, got {t.dtype}: cast at the call site")
    return C.c_void_p(t.data_ptr())


def ptr_i64(t):
    return ptr(t, torch.int64)


def ptr_bf16(t):
    return ptr(t, torch.bfloat16)


def ptr_f64(t):
    return ptr(t, torch.float64)


def stream(device=None):
    """Current torch stream of `device` (default: the current device)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def f32(t, device=None):
    """fp32 contiguous view/copy on `device`."""
    if device is not None:
        t = t.to(device)
    return t.to(torch.float32).contiguous()
