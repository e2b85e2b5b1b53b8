"""ctypes binding of include/dto_b200.h (the C-ABI drop-in boundary).

There is no Python or CPU implementation behind this module: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C dual_threshold_optimization_b200/csrc`) the import
of the library fails loudly, and every compute call fails with DtoError when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdto_b200.so")
CLI_PATH = os.path.join(_HERE, "lib", "dual_threshold_optimization")

OK = 0
ERR_INVALID = -1
ERR_CUDA = -2
ERR_STATE = -3
ERR_PANIC = -4
ERR_UNSUPPORTED = -5
ERR_IO = -6

FLAG_PERMUTED = 0x1
FLAG_TIE_RESOLVED = 0x2  # the host libm settled a tie set the device could not (include/dto_b200.h)
FLAG_HOST_PVALUE = 0x4
FLAG_PATH_FULL = 0x8
FLAG_TIE_MINP = 0x10
FLAG_TIE_OVERLAP = 0x20


class DtoError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dto_b200 error {code}: {message}")
        self.code = code
        self.message = message


class DtoPanic(DtoError):
    """A condition on which the reference panics (message mirrors the reference's)."""


class Record(C.Structure):
    _fields_ = [
        ("rank1", C.c_uint32),
        ("rank2", C.c_uint32),
        ("set1_len", C.c_uint32),
        ("set2_len", C.c_uint32),
        ("intersection_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("population_size", C.c_uint64),
        ("pvalue", C.c_double),
    ]


RECORD_DTYPE = np.dtype(
    [
        ("rank1", "<u4"),
        ("rank2", "<u4"),
        ("set1_len", "<u4"),
        ("set2_len", "<u4"),
        ("intersection_size", "<u4"),
        ("flags", "<u4"),
        ("population_size", "<u8"),
        ("pvalue", "<f8"),
    ]
)
assert RECORD_DTYPE.itemsize == C.sizeof(Record) == 40


class Stats(C.Structure):
    _fields_ = [
        ("tasks_fast", C.c_uint64),
        ("tasks_full", C.c_uint64),
        ("candidates", C.c_uint64),
        ("level2_cells", C.c_uint64),
        ("refined_cells", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("last_scan_kernel_ms", C.c_double),
        ("last_sigma_kernel_ms", C.c_double),
        ("last_scan_launches", C.c_uint64),
        ("last_run_ms", C.c_double),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("lptab_entries", C.c_uint64),
        ("table_cache_hits", C.c_uint64),
        ("tasks_tie_resolved", C.c_uint64),
        ("tie_cells_host", C.c_uint64),
    ]


class FinalResult(C.Structure):
    _fields_ = [
        ("rank1", C.c_uint64),
        ("rank2", C.c_uint64),
        ("set1_len", C.c_uint64),
        ("set2_len", C.c_uint64),
        ("population_size", C.c_uint64),
        ("unpermuted_intersection_size", C.c_uint64),
        ("unpermuted_pvalue", C.c_double),
        ("empirical_pvalue", C.c_double),
        ("fdr", C.c_double),
    ]


# every symbol include/dto_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_recp = C.POINTER(Record)
_strp = C.POINTER(C.c_char_p)

SYMBOLS = {
    "dto_b200_last_error": (C.c_char_p, []),
    "dto_b200_version": (C.c_char_p, []),
    "dto_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "dto_b200_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "dto_b200_destroy": (None, [_vp]),
    "dto_b200_set_problem": (
        C.c_int,
        [_vp, _u32p, C.c_size_t, _u32p, C.c_size_t, _u32p, C.c_size_t, _u32p, C.c_size_t, _i32p, C.c_uint64],
    ),
    "dto_b200_run_unpermuted": (C.c_int, [_vp, _recp]),
    "dto_b200_run_permuted_indices": (C.c_int, [_vp, _u32p, _u32p, C.c_size_t, _recp]),
    "dto_b200_run_permuted_philox": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, _recp, _f64p]),
    "dto_b200_run_permuted_philox_device": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, _vp, _vp]),
    "dto_b200_philox_pairing": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _u32p]),
    "dto_b200_grid_debug": (C.c_int, [_vp, _u32p, _u32p, _u32p, _f64p, _f64p]),
    "dto_b200_hypergeometric_pvalues": (C.c_int, [_vp, _u64p, _u64p, _u64p, _u64p, C.c_size_t, _f64p]),
    "dto_b200_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "dto_b200_reset_stats": (C.c_int, [_vp]),
    "dto_b200_process_totals": (C.c_int, [_u64p]),
    "dto_b200_table_logp": (C.c_int, [_vp, _u32p, _u32p, _u32p, C.c_size_t, _f64p]),
    "dto_b200_last_batch_task_stats": (C.c_int, [_vp, _u32p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dto_b200_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "dto_b200_probe_fp64_tflops": (C.c_int, [_vp, _f64p]),
    "dto_b200_probe_hbm_gbs": (C.c_int, [_vp, _f64p]),
    "dto_b200_ranked_list_from": (C.c_int, [_strp, _u32p, C.c_size_t, C.POINTER(_vp)]),
    "dto_b200_read_ranked_list_csv": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "dto_b200_ranked_list_free": (None, [_vp]),
    "dto_b200_ranked_list_len": (C.c_size_t, [_vp]),
    "dto_b200_ranked_list_num_thresholds": (C.c_size_t, [_vp]),
    "dto_b200_ranked_list_thresholds": (_u32p, [_vp]),
    "dto_b200_ranked_list_ranks": (_u32p, [_vp]),
    "dto_b200_ranked_list_id": (C.c_char_p, [_vp, C.c_size_t]),
    "dto_b200_feature_list_from": (C.c_int, [_strp, C.c_size_t, C.POINTER(_vp)]),
    "dto_b200_read_feature_list": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "dto_b200_feature_list_free": (None, [_vp]),
    "dto_b200_feature_list_len": (C.c_size_t, [_vp]),
    "dto_b200_feature_list_id": (C.c_char_p, [_vp, C.c_size_t]),
    "dto_b200_get_limits": (C.c_int, [C.POINTER(C.c_uint64 * 4)]),
    "dto_b200_compute_population_size": (C.c_int, [_vp, _vp, _vp, _u64p]),
    "dto_b200_load_lists": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "dto_b200_optimize": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, _recp]),
    "dto_b200_run_single_node": (
        C.c_int,
        [_vp, _vp, C.c_uint64, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.c_uint64, _recp],
    ),
    "dto_b200_run_tasks": (
        C.c_int,
        [_vp, _vp, C.c_uint64, _u64p, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.c_uint64, _recp],
    ),
    "dto_b200_nccl_unique_id": (C.c_int, [_vp]),
    "dto_b200_nccl_comm_create": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, C.c_int, C.c_int]),
    "dto_b200_nccl_comm_destroy": (C.c_int, [_vp]),
    "dto_b200_allgather_minima": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t]),
    "dto_b200_run_pairs": (
        C.c_int,
        [C.POINTER(_vp), C.POINTER(_vp), _u64p, C.c_size_t, C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.c_uint64, C.POINTER(FinalResult)],
    ),
    "dto_b200_hypergeometric_pvalue_host": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _f64p]),
    "dto_b200_fdr": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, _f64p]),
    "dto_b200_empirical_pvalue": (C.c_int, [_recp, C.c_size_t, C.POINTER(FinalResult)]),
    "dto_b200_final_result_json": (C.c_int, [C.POINTER(FinalResult), C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
}

_lib = None


def lib():
    """Loads libdto_b200.so (raises if it was not built -- there is nothing to fall back to)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA library first (__graft_entry__.build() or "
                "`make -C dual_threshold_optimization_b200/csrc`). There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def process_totals() -> dict:
    """{launches, h2d_bytes, d2h_bytes} of the whole process since the library was loaded (pooled contexts included)."""
    a = (C.c_uint64 * 3)()
    check(lib().dto_b200_process_totals(a))
    return {"launches": int(a[0]), "h2d_bytes": int(a[1]), "d2h_bytes": int(a[2])}


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = lib().dto_b200_last_error().decode(errors="replace")
    raise (DtoPanic if rc == ERR_PANIC else DtoError)(rc, msg)


def u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def c_strings(items):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode() if isinstance(s, str) else bytes(s)
    return arr
