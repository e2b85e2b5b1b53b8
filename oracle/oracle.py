"""ctypes front-end of the CPU ORACLE (oracle/dto_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package (dual_threshold_optimization_b200) never imports this module.

The oracle restates the reference algorithm (file:line citations live in dto_oracle.c / dto_oracle.h);
its fidelity is pinned by tests/test_oracle_goldens.py against every golden the reference's own tests hold.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdto_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dto_oracle.c")
    hdr = os.path.join(_HERE, "dto_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


class Record(C.Structure):
    _fields_ = [
        ("rank1", C.c_uint32),
        ("rank2", C.c_uint32),
        ("set1_len", C.c_uint32),
        ("set2_len", C.c_uint32),
        ("intersection_size", C.c_uint32),
        ("permuted", C.c_uint32),
        ("population_size", C.c_uint64),
        ("pvalue", C.c_double),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


RECORD_DTYPE = np.dtype(
    [
        ("rank1", "<u4"),
        ("rank2", "<u4"),
        ("set1_len", "<u4"),
        ("set2_len", "<u4"),
        ("intersection_size", "<u4"),
        ("permuted", "<u4"),
        ("population_size", "<u8"),
        ("pvalue", "<f8"),
    ]
)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u32p = C.POINTER(C.c_uint32)
    i32p = C.POINTER(C.c_int32)
    f64p = C.POINTER(C.c_double)
    strp = C.POINTER(C.c_char_p)
    recp = C.POINTER(Record)
    L.oracle_ln_gamma.restype = C.c_double
    L.oracle_ln_gamma.argtypes = [C.c_double]
    L.oracle_ln_factorial.restype = C.c_double
    L.oracle_ln_factorial.argtypes = [C.c_uint64]
    L.oracle_ln_binomial.restype = C.c_double
    L.oracle_ln_binomial.argtypes = [C.c_uint64, C.c_uint64]
    L.oracle_hypergeom_sf.restype = C.c_double
    L.oracle_hypergeom_sf.argtypes = [C.c_uint64] * 4
    L.oracle_hypergeometric_pvalue.restype = C.c_double
    L.oracle_hypergeometric_pvalue.argtypes = [C.c_uint64] * 4
    L.oracle_hypergeometric_pvalue_cached.restype = C.c_double
    L.oracle_hypergeometric_pvalue_cached.argtypes = [f64p] + [C.c_uint64] * 4
    L.oracle_hypergeometric_log_pvalue.restype = C.c_double
    L.oracle_hypergeometric_log_pvalue.argtypes = [f64p] + [C.c_uint64] * 4
    L.oracle_tail_terms.restype = C.c_uint64
    L.oracle_tail_terms.argtypes = [f64p] + [C.c_uint64] * 4
    L.oracle_fill_ln_factorial.restype = None
    L.oracle_fill_ln_factorial.argtypes = [f64p, C.c_uint64]
    L.oracle_stable_sort_by_rank.restype = None
    L.oracle_stable_sort_by_rank.argtypes = [u32p, C.c_size_t, u32p, u32p]
    L.oracle_generate_thresholds.restype = C.c_size_t
    L.oracle_generate_thresholds.argtypes = [u32p, C.c_size_t, u32p, C.c_size_t]
    L.oracle_process_threshold_pairs_faithful.restype = C.c_int
    L.oracle_process_threshold_pairs_faithful.argtypes = [
        strp, u32p, C.c_size_t, u32p, C.c_size_t,
        strp, u32p, C.c_size_t, u32p, C.c_size_t,
        u32p, u32p, C.c_int, C.c_uint64, recp,
    ]
    L.oracle_argmin_tiebreak.restype = C.c_size_t
    L.oracle_argmin_tiebreak.argtypes = [recp, C.c_size_t]
    L.oracle_grid_int.restype = C.c_int
    L.oracle_grid_int.argtypes = [
        u32p, C.c_size_t, u32p, C.c_size_t,
        u32p, C.c_size_t, u32p, C.c_size_t,
        i32p, u32p, u32p, C.c_int, C.c_uint64, f64p,
        u32p, f64p, f64p, recp,
    ]
    L.oracle_grid_int_batch.restype = C.c_int
    L.oracle_grid_int_batch.argtypes = [
        u32p, C.c_size_t, u32p, C.c_size_t,
        u32p, C.c_size_t, u32p, C.c_size_t,
        i32p, u32p, u32p, C.c_size_t, C.c_uint64, f64p, C.c_size_t, recp,
    ]
    L.oracle_fdr.restype = C.c_double
    L.oracle_fdr.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double]
    L.oracle_empirical_pvalue.restype = C.c_double
    L.oracle_empirical_pvalue.argtypes = [f64p, C.c_size_t, C.c_double]
    L.oracle_run_single_node.restype = C.c_int
    L.oracle_run_single_node.argtypes = [
        strp, u32p, C.c_size_t, strp, u32p, C.c_size_t, i32p, C.c_uint64,
        C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, recp,
    ]
    L.oracle_run_single_node_sampled.restype = C.c_int
    L.oracle_run_single_node_sampled.argtypes = [
        strp, u32p, C.c_size_t, strp, u32p, C.c_size_t, i32p, C.c_uint64,
        C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.c_size_t, recp,
    ]
    L.oracle_grid_tail_terms.restype = None
    L.oracle_grid_tail_terms.argtypes = [u32p, u32p, C.c_size_t, u32p, C.c_size_t, C.c_uint64, f64p,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.oracle_shuffle.restype = None
    L.oracle_shuffle.argtypes = [u32p, C.c_size_t, C.c_uint64]
    _lib = L
    return L


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a: Optional[np.ndarray], ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def _strs(ids: Sequence[str]):
    arr = (C.c_char_p * max(len(ids), 1))()
    for i, s in enumerate(ids):
        arr[i] = s.encode()
    return arr


# ---------------------------------------------------------------------------------------------------
# scalar statistics
# ---------------------------------------------------------------------------------------------------
def ln_factorial(x: int) -> float:
    return lib().oracle_ln_factorial(x)


def ln_factorial_table(N: int) -> np.ndarray:
    lf = np.empty(N + 1, dtype=np.float64)
    lib().oracle_fill_ln_factorial(_p(lf, C.c_double), N)
    return lf


def hypergeometric_pvalue(N: int, K: int, n: int, k: int) -> float:
    """stat_operations/hypergeometric_pvalue.rs:33-50 (uncached, full ascending tail)."""
    return lib().oracle_hypergeometric_pvalue(N, K, n, k)


def hypergeometric_pvalue_cached(lf: np.ndarray, N: int, K: int, n: int, k: int) -> float:
    return lib().oracle_hypergeometric_pvalue_cached(_p(lf, C.c_double), N, K, n, k)


def hypergeometric_log_pvalue(lf: np.ndarray, N: int, K: int, n: int, k: int) -> float:
    return lib().oracle_hypergeometric_log_pvalue(_p(lf, C.c_double), N, K, n, k)


def tail_terms(lf: np.ndarray, N: int, K: int, n: int, k: int) -> int:
    return lib().oracle_tail_terms(_p(lf, C.c_double), N, K, n, k)


def fdr(b: int, r: int, k: int, N: int, sensitivity: float = 0.8) -> float:
    return lib().oracle_fdr(b, r, k, N, sensitivity)


def empirical_pvalue(permuted_minp, unpermuted_p: float) -> float:
    a = np.ascontiguousarray(permuted_minp, dtype=np.float64)
    return lib().oracle_empirical_pvalue(_p(a, C.c_double), a.size, unpermuted_p)


def shuffle(n: int, seed: int) -> np.ndarray:
    idx = np.empty(n, dtype=np.uint32)
    lib().oracle_shuffle(_p(idx, C.c_uint32), n, seed)
    return idx


# ---------------------------------------------------------------------------------------------------
# collections
# ---------------------------------------------------------------------------------------------------
@dataclass
class OracleRankedList:
    """RankedFeatureList::from (collections/ranked.rs:176-191): stable sort by rank + thresholds."""

    ids: list
    ranks: np.ndarray  # sorted ascending, uint32
    thresholds: np.ndarray  # uint32

    @staticmethod
    def make(ids: Sequence[str], ranks) -> "OracleRankedList":
        ranks = _u32(ranks)
        if len(ids) != ranks.size:
            raise ValueError("Genes and ranks must have the same length.")
        n = ranks.size
        sr = np.empty(n, dtype=np.uint32)
        order = np.empty(n, dtype=np.uint32)
        lib().oracle_stable_sort_by_rank(_p(ranks, C.c_uint32), n, _p(sr, C.c_uint32), _p(order, C.c_uint32))
        T = lib().oracle_generate_thresholds(_p(sr, C.c_uint32), n, None, 0)
        thr = np.empty(T, dtype=np.uint32)
        lib().oracle_generate_thresholds(_p(sr, C.c_uint32), n, _p(thr, C.c_uint32), T)
        return OracleRankedList([ids[i] for i in order], sr, thr)


def slot_map(l1: OracleRankedList, l2: OracleRankedList) -> np.ndarray:
    pos2 = {g: j for j, g in enumerate(l2.ids)}
    if len(pos2) != len(l2.ids) or len(set(l1.ids)) != len(l1.ids):
        raise ValueError("integer mode does not support duplicate ids")
    return np.array([pos2.get(g, -1) for g in l1.ids], dtype=np.int32)


def compute_population_size(l1: OracleRankedList, l2: OracleRankedList, background: Optional[Sequence[str]]) -> int:
    """dto/compute_population_size.rs:66-104 (panics -> ValueError)."""
    if background is None:
        s1, s2 = set(l1.ids), set(l2.ids)
        # FeatureList::intersect keeps list-1 items present in list 2 (feature_list.rs) -> len of that
        inter = sum(1 for g in l1.ids if g in s2)
        if inter != len(l1.ids) or inter != len(l2.ids):
            raise ValueError("If no background is provided, the feature lists must have identical genes.")
        return inter
    bg = set(background)
    for lst, name in ((l1, "first"), (l2, "second")):
        diff = [g for g in lst.ids if g not in bg]
        if diff:
            raise ValueError(f"The following genes in the {name} ranked feature list are not in the background: {diff}")
    return len(background)


# ---------------------------------------------------------------------------------------------------
# grids
# ---------------------------------------------------------------------------------------------------
def process_threshold_pairs_faithful(l1, l2, population, perm1=None, perm2=None, permuted=None) -> np.ndarray:
    """dto/process_threshold_pairs.rs:71-131 in the reference's own data structures. Returns T1*T2 records."""
    T1, T2 = l1.thresholds.size, l2.thresholds.size
    out = np.zeros(T1 * T2, dtype=RECORD_DTYPE)
    p1 = None if perm1 is None else _u32(perm1)
    p2 = None if perm2 is None else _u32(perm2)
    flag = (perm1 is not None) if permuted is None else permuted
    ids1, ids2 = _strs(l1.ids), _strs(l2.ids)
    rc = lib().oracle_process_threshold_pairs_faithful(
        ids1, _p(l1.ranks, C.c_uint32), len(l1.ids), _p(l1.thresholds, C.c_uint32), T1,
        ids2, _p(l2.ranks, C.c_uint32), len(l2.ids), _p(l2.thresholds, C.c_uint32), T2,
        _p(p1, C.c_uint32), _p(p2, C.c_uint32), int(bool(flag)), population,
        out.ctypes.data_as(C.POINTER(Record)),
    )
    if rc != 0:
        raise ValueError("Failed to create hypergeometric distribution")
    return out


def argmin_tiebreak(records: np.ndarray) -> np.void:
    """dto/optimize_main.rs:73-116."""
    if records.size == 0:
        raise ValueError("empty grid (optimize_main.rs:116 unwrap)")
    i = lib().oracle_argmin_tiebreak(records.ctypes.data_as(C.POINTER(Record)), records.size)
    return records[i]


def optimize_faithful(l1, l2, population, perm1=None, perm2=None):
    return argmin_tiebreak(process_threshold_pairs_faithful(l1, l2, population, perm1, perm2))


@dataclass
class GridResult:
    overlap: Optional[np.ndarray]
    p: Optional[np.ndarray]
    logp: Optional[np.ndarray]
    best: dict


def grid_int(l1, l2, population, slot2_of_1=None, perm1=None, perm2=None, lf=None,
             want_overlap=True, want_p=True, want_logp=False, permuted=None) -> GridResult:
    """Integer-id form of the same grid (histogram + 2-D prefix sum), cross-checked vs the faithful form."""
    T1, T2 = l1.thresholds.size, l2.thresholds.size
    if slot2_of_1 is None:
        slot2_of_1 = slot_map(l1, l2)
    if lf is None:
        lf = ln_factorial_table(population)
    ov = np.zeros((T1, T2), dtype=np.uint32) if want_overlap else None
    pp = np.zeros((T1, T2), dtype=np.float64) if want_p else None
    lp = np.zeros((T1, T2), dtype=np.float64) if want_logp else None
    best = Record()
    p1 = None if perm1 is None else _u32(perm1)
    p2 = None if perm2 is None else _u32(perm2)
    flag = (perm1 is not None) if permuted is None else permuted
    rc = lib().oracle_grid_int(
        _p(l1.ranks, C.c_uint32), len(l1.ids), _p(l1.thresholds, C.c_uint32), T1,
        _p(l2.ranks, C.c_uint32), len(l2.ids), _p(l2.thresholds, C.c_uint32), T2,
        _p(np.ascontiguousarray(slot2_of_1, dtype=np.int32), C.c_int32),
        _p(p1, C.c_uint32), _p(p2, C.c_uint32), int(bool(flag)), population, _p(lf, C.c_double),
        _p(ov, C.c_uint32), _p(pp, C.c_double), _p(lp, C.c_double), C.byref(best),
    )
    if rc == -1:
        raise ValueError("Failed to create hypergeometric distribution")
    if rc != 0:
        raise ValueError("empty threshold list")
    return GridResult(ov, pp, lp, best.as_dict())


def best_batch(l1, l2, population, perm1, perm2, slot2_of_1=None, lf=None, num_threads=None) -> np.ndarray:
    """Best record (optimize_main.rs:73-116) of every permuted task given by the index rows perm1[t], perm2[t]: the
    integer grid on `num_threads` OS threads (default: all cores).  For parity runs over thousands of permutations."""
    import os

    p1 = np.ascontiguousarray(perm1, dtype=np.uint32)
    p2 = np.ascontiguousarray(perm2, dtype=np.uint32)
    assert p1.ndim == 2 and p2.ndim == 2 and p1.shape[0] == p2.shape[0]
    assert p1.shape[1] == len(l1.ids) and p2.shape[1] == len(l2.ids)
    if slot2_of_1 is None:
        slot2_of_1 = slot_map(l1, l2)
    if lf is None:
        lf = ln_factorial_table(population)
    out = np.zeros(p1.shape[0], dtype=RECORD_DTYPE)
    rc = lib().oracle_grid_int_batch(
        _p(l1.ranks, C.c_uint32), len(l1.ids), _p(l1.thresholds, C.c_uint32), l1.thresholds.size,
        _p(l2.ranks, C.c_uint32), len(l2.ids), _p(l2.thresholds, C.c_uint32), l2.thresholds.size,
        _p(np.ascontiguousarray(slot2_of_1, dtype=np.int32), C.c_int32), _p(p1, C.c_uint32), _p(p2, C.c_uint32),
        p1.shape[0], population, _p(lf, C.c_double), num_threads or (os.cpu_count() or 1),
        out.ctypes.data_as(C.POINTER(Record)),
    )
    if rc != 0:
        raise ValueError(f"oracle_grid_int_batch failed rc={rc}")
    return out


def run_single_node(l1, l2, population, task_permute, num_threads, seed=0, mode=0, slot2_of_1=None, row_stride=1) -> np.ndarray:
    """run/single_node.rs:83-137 (static chunks on OS threads); mode 0 = reference-faithful, 1 = integer.
    row_stride > 1 (faithful mode only) evaluates every row_stride-th t1 row: a bounded timing sample."""
    tp = np.ascontiguousarray(task_permute, dtype=np.uint8)
    out = np.zeros(tp.size, dtype=RECORD_DTYPE)
    if slot2_of_1 is None and mode != 0:
        slot2_of_1 = slot_map(l1, l2)
    sm = None if slot2_of_1 is None else np.ascontiguousarray(slot2_of_1, dtype=np.int32)
    ids1, ids2 = _strs(l1.ids), _strs(l2.ids)
    rc = lib().oracle_run_single_node_sampled(
        ids1, _p(l1.ranks, C.c_uint32), len(l1.ids), ids2, _p(l2.ranks, C.c_uint32), len(l2.ids),
        _p(sm, C.c_int32), population, _p(tp, C.c_uint8), tp.size, num_threads, seed, mode, row_stride,
        out.ctypes.data_as(C.POINTER(Record)),
    )
    if rc != 0:
        raise ValueError(f"oracle_run_single_node failed rc={rc}")
    return out


def grid_tail_terms(l1, l2, population, overlap, lf=None):
    """(sum of converged tail lengths R, number of cells the reference actually evaluates) -- SURVEY 8(d)."""
    if lf is None:
        lf = ln_factorial_table(population)
    c1 = np.searchsorted(l1.ranks, l1.thresholds, side="right").astype(np.uint32)
    c2 = np.searchsorted(l2.ranks, l2.thresholds, side="right").astype(np.uint32)
    ov = np.ascontiguousarray(overlap, dtype=np.uint32)
    terms, cells = C.c_uint64(), C.c_uint64()
    lib().oracle_grid_tail_terms(_p(ov, C.c_uint32), _p(c1, C.c_uint32), c1.size, _p(c2, C.c_uint32), c2.size,
                                 population, _p(lf, C.c_double), C.byref(terms), C.byref(cells))
    return terms.value, cells.value


def final_json(results: np.ndarray) -> dict:
    """stat_operations/empirical_pvalue.rs:109-187 on a vector of Best records."""
    unperm = results[results["permuted"] == 0]
    perm = results[results["permuted"] != 0]
    if unperm.size == 0:
        raise ValueError("No unpermuted result found in the provided results.")
    u = unperm[0]
    out = {
        "rank1": int(u["rank1"]),
        "rank2": int(u["rank2"]),
        "set1_len": int(u["set1_len"]),
        "set2_len": int(u["set2_len"]),
        "population_size": int(u["population_size"]),
        "unpermuted_intersection_size": int(u["intersection_size"]),
        "unpermuted_pvalue": float(u["pvalue"]),
        "empirical_pvalue": empirical_pvalue(perm["pvalue"], float(u["pvalue"])),
        "fdr": fdr(int(u["set1_len"]), int(u["set2_len"]), int(u["intersection_size"]), int(u["population_size"]), 0.8),
    }
    return out
