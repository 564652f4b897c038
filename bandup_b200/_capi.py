"""ctypes binding of include/bandup_gpu.h -- the only way Python reaches the kernels.

Loading fails loudly when libbandup_gpu.so is missing: there is no CPU or
PyTorch fallback for this path anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libbandup_gpu.so"
HEADER = PKG.parent / "include" / "bandup_gpu.h"

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)

OK = 0
ERR_NAMES = {}


class KptDesc(C.Structure):
    """struct bandup_gpu_kpt_desc (include/bandup_gpu.h)."""
    _fields_ = [("n_spinor", C.c_int32), ("n_pw", C.c_int32), ("n_bands", C.c_int32), ("layout", C.c_int32),
                ("n_fold", C.c_int32), ("reserved", C.c_int32), ("band_stride_bytes", C.c_int64),
                ("d_coeffs", C.c_void_p), ("d_g_frac", C.c_void_p), ("d_band_energies", C.c_void_p),
                ("g0", C.c_void_p), ("d_weights", C.c_void_p), ("d_dN", C.c_void_p),
                ("d_n_selected", C.c_void_p), ("d_norms", C.c_void_p)]


class BandupGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"bandup_gpu error {code} ({ERR_NAMES.get(code, '?')}): {message}")
        self.code = code


def header_constants() -> dict:
    """#define NAME <int> pairs of the public header."""
    out = {}
    for m in re.finditer(r"^#define\s+(BANDUP_GPU_\w+)\s+(-?\d+)\s*(?:/\*.*)?$", HEADER.read_text(), re.M):
        out[m.group(1)] = int(m.group(2))
    return out


def header_symbols() -> list:
    """Names of every function the public header declares."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(bandup_gpu_\w+)\s*\(", text)))


CONST = header_constants()
for _k, _v in CONST.items():
    if _k.startswith("BANDUP_GPU_ERR_") or _k == "BANDUP_GPU_OK":
        ERR_NAMES[_v] = _k

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m bandup_b200.build` "
            "(needs nvcc). bandup_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    sig = {
        "bandup_gpu_abi_version": (C.c_int, []),
        "bandup_gpu_init": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
        "bandup_gpu_finalize": (C.c_int, [C.c_int]),
        "bandup_gpu_last_error": (C.c_char_p, [C.c_int]),
        "bandup_gpu_pinned_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
        "bandup_gpu_pinned_free": (C.c_int, [C.c_void_p]),
        "bandup_gpu_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
        "bandup_gpu_host_unregister": (C.c_int, [C.c_void_p]),
        "bandup_gpu_set_cells": (C.c_int, [C.c_int, c_f64p, c_f64p, C.c_double, c_i32p]),
        "bandup_gpu_set_grid": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_i32p]),
        "bandup_gpu_get_grid": (C.c_int, [C.c_int, c_f64p]),
        "bandup_gpu_folding_g0": (C.c_int, [C.c_int, c_f64p, C.c_int, c_i32p, c_f64p]),
        "bandup_gpu_fold_map": (C.c_int, [C.c_int, C.c_int, c_f64p, C.c_int, c_i32p, c_f64p, C.c_int, c_i32p, c_i32p,
                                          c_f64p]),
        "bandup_gpu_submit_kpt": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                            C.c_void_p, c_i32p, c_f64p, C.c_int, c_i32p]),
        "bandup_gpu_submit_kpt_vasp": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                                 c_f64p, C.c_double, C.c_double, c_f64p, C.c_int, c_i32p,
                                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "bandup_gpu_fetch_gvecs": (C.c_int, [C.c_int, C.c_int, c_i32p, C.c_int64]),
        "bandup_gpu_fetch_kpt": (C.c_int, [C.c_int, C.c_int, c_f64p, c_f64p, c_i32p, c_f64p, c_i32p, C.c_int64]),
        "bandup_gpu_pipeline_depth": (C.c_int, [C.c_int]),
        "bandup_gpu_rho_count": (C.c_int, [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64)]),
        "bandup_gpu_fetch_rho": (C.c_int, [C.c_int, C.c_int, C.c_int64, c_i32p, c_i32p, c_i32p, c_i32p,
                                           C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "bandup_gpu_orb_contrib_matrix": (C.c_int, [C.c_int, C.c_int, C.c_int, c_f64p, c_f64p, c_u8p, c_f64p,
                                                    C.c_double, C.c_double, c_f64p]),
        "bandup_gpu_set_operator": (C.c_int, [C.c_int, C.c_int, c_f64p]),
        "bandup_gpu_pauli_vectors": (C.c_int, [C.c_int, C.c_int, c_f64p]),
        "bandup_gpu_set_perform_unfold": (C.c_int, [C.c_int, C.c_int]),
        "bandup_gpu_spectral_function": (C.c_int, [C.c_int, C.c_int, C.c_double, c_f64p]),
        "bandup_gpu_unfold_operator": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, c_f64p]),
        "bandup_gpu_symm_average": (C.c_int, [C.c_int, C.c_int, C.c_int64, c_f64p, c_f64p, c_f64p]),
        "bandup_gpu_symm_average_rho": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f64p, c_f64p,
                                                  C.c_double, C.POINTER(C.c_int64), c_i32p, c_i32p, c_i32p, c_i32p,
                                                  C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64,
                                                  C.POINTER(C.c_int64), c_i32p, c_i32p, c_i32p, c_i32p,
                                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "bandup_gpu_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "bandup_gpu_unfold_device": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_i32p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_size_t]),
        "bandup_gpu_batch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.POINTER(KptDesc)]),
        "bandup_gpu_unfold_device_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.POINTER(KptDesc),
                                                     C.c_void_p, C.c_size_t]),
        "bandup_gpu_comm_unique_id": (C.c_int, [C.c_void_p]),
        "bandup_gpu_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p]),
        "bandup_gpu_comm_finalize": (C.c_int, [C.c_int]),
        "bandup_gpu_gather_dN": (C.c_int, [C.c_int, c_f64p, C.POINTER(C.c_int64), C.c_int64, c_f64p,
                                           C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)]),
        "bandup_gpu_gather_rows": (C.c_int, [C.c_int, C.c_int, c_f64p, C.POINTER(C.c_int64), C.c_int64, c_f64p,
                                             C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)]),
        "bandup_gpu_allgather_device": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
        "bandup_gpu_set_profiling": (C.c_int, [C.c_int, C.c_int]),
        "bandup_gpu_set_profiling_mask": (C.c_int, [C.c_int, C.c_uint]),
        "bandup_gpu_kernel_time": (C.c_int, [C.c_int, C.c_int, c_f64p, C.POINTER(C.c_int64)]),
        "bandup_gpu_reset_kernel_times": (C.c_int, [C.c_int]),
        "bandup_gpu_launch_count": (C.c_int64, [C.c_int]),
        "bandup_gpu_set_tuning": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
        "bandup_gpu_set_option": (C.c_int, [C.c_int, C.c_int, C.c_int]),
        "bandup_gpu_synth_fill_device": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                                   C.c_int64, C.c_uint64, C.c_uint64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.bandup_gpu_abi_version() != CONST["BANDUP_GPU_ABI_VERSION"]:
        raise ImportError("libbandup_gpu.so ABI version does not match include/bandup_gpu.h; rebuild")
    _lib = lib
    return lib


def check(handle: int, rc: int) -> None:
    if rc != OK:
        msg = load().bandup_gpu_last_error(handle)
        raise BandupGpuError(rc, msg.decode() if msg else "")
