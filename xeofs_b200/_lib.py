"""ctypes binding of libxeofs_b200.so (the C-ABI declared in include/xeofs_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, this module
raises.  Device buffers are torch CUDA tensors; only their ``data_ptr()`` crosses the boundary.
"""
from __future__ import annotations

import ctypes as C
import os


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxeofs_b200.so")

ALGO_AUTO, ALGO_SIMT, ALGO_TF32X1, ALGO_TF32X3, ALGO_AUTO_FAST, ALGO_TF32X2, ALGO_TF32X1R, ALGO_TF32X1F = 0, 1, 2, 3, 4, 5, 6, 7
ALGO_NAMES = {"auto": ALGO_AUTO, "simt": ALGO_SIMT, "tf32x1": ALGO_TF32X1, "tf32x3": ALGO_TF32X3, "tf32x2": ALGO_TF32X2, "tf32x1r": ALGO_TF32X1R}
ALGO_FLAG_NO_NAN = 0x100
F_CENTER, F_STANDARDIZE = 1, 2
E_INVALID, E_CUDA, E_WORKSPACE, E_UNSUPPORTED = -1, -2, -3, -4

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> (restype, argtypes); every symbol include/xeofs_b200.h declares
SIGNATURES = {
    "xeofs_b200_version": (_int, []),
    "xeofs_b200_last_error": (C.c_char_p, []),
    "xeofs_b200_has_tcgen05": (_int, []),
    "xeofs_b200_col_stats": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "xeofs_b200_scaling_finalize": (_int, [_i64, _p, _p, _p, _p, _p, _int, _p, _p, _p, _p, _p, _p, _p, _p]),
    "xeofs_b200_project_workspace_bytes": (_i64, [_i64, _i64, _i64, _int]),
    "xeofs_b200_project_S": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _p, _i64, _int, _p]),
    "xeofs_b200_project_T": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _p, _i64, _int, _p]),
    "xeofs_b200_project_S_stats": (_int, [_p, _i64, _i64, _i64, _p, _int, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                          _i64, _p, _i64, _p]),
    "xeofs_b200_h16_scales": (_int, [_p, _p, _i64, _p, _p, _p]),
    "xeofs_b200_project_T_h16copy": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i64, _i64, _p, _i64, _p, _i64, _int, _p, _p,
                                            _i64, _p]),
    "xeofs_b200_project_S16": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i64, _i64, _p, _i64, _p, _i64, _p]),
    "xeofs_b200_project_T16": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i64, _i64, _p, _i64, _p, _i64, _p]),
    "xeofs_b200_project_S_stats_h16copy": (_int, [_p, _i64, _i64, _i64, _p, _int, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p,
                                                  _p, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _p]),
    "xeofs_b200_round_tf32": (_int, [_p, _i64, _i64, _i64, _p]),
    "xeofs_b200_gram": (_int, [_p, _i64, _i64, _i64, _int, _p, _int, _p]),
    "xeofs_b200_chol_inv": (_int, [_p, _i64, _p, _p, _p]),
    "xeofs_b200_apply": (_int, [_p, _i64, _i64, _i64, _int, _p, _i64, _i64, _p, _p, _i64, _p]),
    "xeofs_b200_sym_eig": (_int, [_p, _i64, _p, _p, _p, _p, _p]),
    "xeofs_b200_row_minmax": (_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "xeofs_b200_finish_components": (_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "xeofs_b200_varimax_accumulate": (_int, [_p, _i64, _i64, _i64, _p, _p, C.c_double, _p, _p, _p, _p, _int, _p]),
    "xeofs_b200_varimax_workspace_bytes": (_i64, [_i64, _i64]),
    "xeofs_b200_varimax_pack_bytes": (_i64, [_i64, _i64]),
    "xeofs_b200_varimax_pack": (_int, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "xeofs_b200_varimax_sweep": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _int, _int, _p, _i64, _p]),
    "xeofs_b200_varimax_update_workspace_bytes": (_i64, [_i64]),
    "xeofs_b200_varimax_update": (_int, [_p, _p, _p, C.c_double, _i64, _p, _p, _p, C.c_double, _p, _i64, _p]),
    "xeofs_b200_col_norms": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i64, _p]),
    "xeofs_b200_scaled_rows": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _p, _i64, _p]),
    "xeofs_b200_dgemm": (_int, [_int, _int, _i64, _i64, _i64, C.c_double, _p, _i64, _p, _i64, C.c_double, _p, _i64, _p]),
    "xeofs_b200_gram_wide": (_int, [_p, _i64, _i64, _i64, _int, _p, _p]),
    "xeofs_b200_sym_eig_wide_workspace_bytes": (_i64, [_i64]),
    "xeofs_b200_sym_eig_wide": (_int, [_p, _i64, _p, _p, _p, _i64, _p, _int, _p]),
    "xeofs_b200_materialize": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _int, _p, _i64, _p]),
    "xeofs_b200_materialize_bf16": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _p, _i64, _p]),
    "xeofs_b200_gram_rows_bf16_workspace_bytes": (_i64, [_i64, _i64]),
    "xeofs_b200_gram_rows_bf16": (_int, [_p, _i64, _i64, _i64, _p, _i64, _p, _i64, _p]),
    "xeofs_b200_reconstruct": (_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64, _p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with ./build.sh (or __graft_entry__.build()). "
                "xeofs_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class XeofsB200Error(RuntimeError):
    pass


def check(rc, what=""):
    """Map the C-ABI error classes onto the exception types the reference raises at this boundary."""
    if rc == 0:
        return
    msg = load().xeofs_b200_last_error().decode("utf-8", "replace")
    msg = f"{what}: {msg}" if what else msg
    if rc in (E_INVALID, E_WORKSPACE):
        raise ValueError(msg)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise XeofsB200Error(msg)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def lpad(l: int) -> int:
    return (int(l) + 15) // 16 * 16
