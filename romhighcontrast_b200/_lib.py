"""ctypes binding of libromhc.so (the C ABI declared in include/romhc.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
product path raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C romhighcontrast_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ROMHC_LIB_PATH") or os.path.join(_HERE, "libromhc.so")   # (the override serves A/B probes of two builds)

ROMHC_OK, ERR_ARG, ERR_CUDA, ERR_NUMERIC, ERR_NOTCONVERGED = 0, 1, 2, 3, 4


class RomhcError(RuntimeError):
    pass


_lib = None

_vp, _i, _i64, _dbl, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); the list must stay in sync with include/romhc.h (tests/test_abi.py checks it)
SIGNATURES = {
    "romhc_version": (_i, []),
    "romhc_last_error": (C.c_char_p, []),
    "romhc_create": (_i, [_i, _i, _i, _i, C.POINTER(_vp)]),
    "romhc_destroy": (_i, [_vp]),
    "romhc_set_option": (_i, [_vp, C.c_char_p, _dbl]),
    "romhc_get_info": (_i, [_vp, C.POINTER(_i64)]),
    "romhc_get_profile": (_i, [_vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    "romhc_check_guards": (_i, [_vp, C.POINTER(_i64)]),
    "romhc_launch_count": (_i64, []),
    "romhc_malloc": (_i, [C.POINTER(_vp), _sz]),
    "romhc_free": (_i, [_vp]),
    "romhc_malloc_host": (_i, [C.POINTER(_vp), _sz]),
    "romhc_free_host": (_i, [_vp]),
    "romhc_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "romhc_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "romhc_memset": (_i, [_vp, _i, _sz, _vp]),
    "romhc_stream_sync": (_i, [_vp]),
    "romhc_pack": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "romhc_unpack": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "romhc_pack_host": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "romhc_unpack_host": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "romhc_apply": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "romhc_energy_norm": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "romhc_l2_norm": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "romhc_error_norm": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "romhc_solve": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(_i64)]),
    "romhc_solve_rhs": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(_i64)]),
    "romhc_precond": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "romhc_project_operators": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "romhc_reduced_solve": (_i, [_vp, _i, _vp, _vp, _i, _i, _i64, _vp, _vp, _vp]),
    "romhc_gemm_nt": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i, _vp]),
    "romhc_gemm_nn": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp]),
    "romhc_gemm_tn": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp]),
    "romhc_tsqr_r": (_i, [_vp, _i64, _i, _i64, _vp, _vp]),
    "romhc_column_mean": (_i, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "romhc_center_rows": (_i, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "romhc_evaluate": (_i, [_vp, _vp, _i, _vp, _i64, _vp, _vp]),
    "romhc_interp_weights": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "romhc_row_norms": (_i, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "romhc_row_dots": (_i, [_vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp]),
    "romhc_estimator": (_i, [_vp, _i64, _i, _vp, _i, _i, _vp, _vp]),
    "romhc_argmax": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "romhc_poly_features": (_i, [_vp, _i64, _i, _i64, _vp, _i, _i, _vp, _i64, _vp]),
    "romhc_generate_solutions_host": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "romhc_reduced_galerkin_host": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
}


def load():
    """Load libromhc.so (once) and declare every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RomhcError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Run `make -C romhighcontrast_b200/csrc` (needs nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc == ROMHC_OK:
        return
    msg = load().romhc_last_error().decode("utf-8", "replace")
    if rc == ERR_NUMERIC:
        raise np.linalg.LinAlgError(msg or "Matrix is singular.")   # reference: LAPACK failure -> LinAlgError
    raise RomhcError(f"libromhc error {rc}: {msg}")


def call(name: str, *args):
    check(getattr(load(), name)(*args))


def launch_count() -> int:
    return int(load().romhc_launch_count())
