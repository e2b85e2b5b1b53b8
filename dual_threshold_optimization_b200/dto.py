"""Mirror of the reference's `dto` module (src/dto/*.rs)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Union

import numpy as np

from . import _capi as capi
from .collections import FeatureList, PermutedRankedFeatureList, RankedFeatureList
from .engine import Engine


class FeatureSets:
    """FeatureSets::Both of the debug path (src/dto/results_objects.rs:8-12, process_threshold_pairs.rs:111-115): the two
    thresholded feature sets of one cell, rebuilt on the host from the (permuted) lists -- the device only ever sees
    integer slots.  Holds views, not copies: a debug grid has T1 x T2 of these."""

    __slots__ = ("_ids1", "_n1", "_ids2", "_n2")

    def __init__(self, ids1, n1, ids2, n2):
        self._ids1, self._n1, self._ids2, self._n2 = ids1, n1, ids2, n2

    def set1(self):
        from .collections import Feature

        return [Feature(g) for g in self._ids1[: self._n1]]

    def set2(self):
        from .collections import Feature

        return [Feature(g) for g in self._ids2[: self._n2]]

    def both(self):
        return self.set1(), self.set2()


@dataclass
class OptimizationResultRecord:
    """src/dto/results_objects.rs:22-32 (feature_sets is FeatureSets::None on the non-debug path)."""

    rank1: int
    rank2: int
    set1_len: int
    set2_len: int
    population_size: int
    intersection_size: int
    pvalue: float
    permuted: bool
    tie_resolved: bool = False  # DTO_B200_FLAG_TIE_RESOLVED: the host libm settled the optimum among ulp-close cells
    feature_sets: Optional[FeatureSets] = None  # FeatureSets::Both when debug=True, FeatureSets::None otherwise

    @staticmethod
    def from_np(r) -> "OptimizationResultRecord":
        return OptimizationResultRecord(
            int(r["rank1"]), int(r["rank2"]), int(r["set1_len"]), int(r["set2_len"]), int(r["population_size"]),
            int(r["intersection_size"]), float(r["pvalue"]), bool(int(r["flags"]) & capi.FLAG_PERMUTED),
            bool(int(r["flags"]) & capi.FLAG_TIE_RESOLVED),
        )


def records_to_array(results) -> np.ndarray:
    if isinstance(results, np.ndarray):
        return np.ascontiguousarray(results, dtype=capi.RECORD_DTYPE)
    arr = np.zeros(len(results), dtype=capi.RECORD_DTYPE)
    for i, r in enumerate(results):
        arr[i] = (r.rank1, r.rank2, r.set1_len, r.set2_len, r.intersection_size,
                  (capi.FLAG_PERMUTED if r.permuted else 0) | (capi.FLAG_TIE_RESOLVED if r.tie_resolved else 0),
                  r.population_size, r.pvalue)
    return arr


def compute_population_size(l1: RankedFeatureList, l2: RankedFeatureList, background: Optional[FeatureList] = None) -> int:
    """src/dto/compute_population_size.rs:66-104 (the reference's panics raise DtoPanic with the same text)."""
    out = C.c_uint64()
    capi.check(capi.lib().dto_b200_compute_population_size(l1.handle, l2.handle, background.handle if background is not None else None, C.byref(out)))
    return out.value


_engine: Optional[Engine] = None


def _default_engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


def process_threshold_pairs(l1: RankedFeatureList, l2: RankedFeatureList, use_permutation: Union[bool, tuple],
                            population_size: int, debug: bool = False, engine: Optional[Engine] = None) -> List[OptimizationResultRecord]:
    """src/dto/process_threshold_pairs.rs:71-131: one record per (t1, t2), row-major.  `use_permutation` may be a
    (PermutedRankedFeatureList, PermutedRankedFeatureList) pair to fix the indices (parity mode) or True to draw
    them with numpy."""
    eng = engine or _default_engine()
    eng.load_lists(l1, l2, population_size)
    perm1 = perm2 = None
    permuted = bool(use_permutation)
    if isinstance(use_permutation, tuple):
        perm1, perm2 = use_permutation[0].indices, use_permutation[1].indices
    elif use_permutation:
        perm1 = PermutedRankedFeatureList(l1).indices
        perm2 = PermutedRankedFeatureList(l2).indices
    ov, pv, _ = eng.grid_debug(perm1, perm2, want_p=True)
    t1, t2 = l1.thresholds(), l2.thresholds()
    r1, r2 = l1.ranks(), l2.ranks()
    c1 = np.searchsorted(r1, t1, side="right")
    c2 = np.searchsorted(r2, t2, side="right")
    ids1 = ids2 = None
    if debug:  # feature sets in (permuted) position order: position j holds the gene of slot indices[j] (permuted.rs:95-99)
        ids1, ids2 = l1.ids(), l2.ids()
        if perm1 is not None:
            ids1 = [ids1[k] for k in perm1]
            ids2 = [ids2[k] for k in perm2]
    out = []
    for i in range(t1.size):
        for j in range(t2.size):
            out.append(OptimizationResultRecord(int(t1[i]), int(t2[j]), int(c1[i]), int(c2[j]), int(population_size),
                                                int(ov[i, j]), float(pv[i, j]), permuted,
                                                feature_sets=FeatureSets(ids1, int(c1[i]), ids2, int(c2[j])) if debug else None))
    return out


def optimize(l1: RankedFeatureList, l2: RankedFeatureList, permute: Union[bool, tuple], population_size: int,
             debug: bool = False, engine: Optional[Engine] = None, seed: int = 0, perm_id: int = 0):
    """src/dto/optimize_main.rs:53-118.  debug=True returns every record (OptimizationResult::Debug); otherwise
    the single best record after the reference's reduction: min p, then max intersection, then min (rank1, rank2)."""
    eng = engine or _default_engine()
    if debug:
        return process_threshold_pairs(l1, l2, permute, population_size, True, eng)
    if isinstance(permute, tuple):
        eng.load_lists(l1, l2, population_size)
        rec = eng.run_permuted_indices(permute[0].indices[None, :], permute[1].indices[None, :])[0]
        return OptimizationResultRecord.from_np(rec)
    rec = np.zeros(1, dtype=capi.RECORD_DTYPE)
    capi.check(capi.lib().dto_b200_optimize(eng.ctx, l1.handle, l2.handle, int(bool(permute)), int(population_size),
                                            int(seed), int(perm_id), rec.ctypes.data_as(C.POINTER(capi.Record))))
    return OptimizationResultRecord.from_np(rec[0])
