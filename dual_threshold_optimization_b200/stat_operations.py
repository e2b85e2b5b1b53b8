"""Mirror of the reference's `stat_operations` module (src/stat_operations/*.rs)."""
from __future__ import annotations

import ctypes as C
import json
from typing import Optional

import numpy as np

from . import _capi as capi
from .engine import Engine

_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def hypergeometric_pvalue(population_size: int, successes_in_population: int, sample_size: int, observed_overlap: int,
                          engine: Optional[Engine] = None) -> float:
    """src/stat_operations/hypergeometric_pvalue.rs:33-50, evaluated ON THE GPU in statrs' operation order."""
    eng = engine or default_engine()
    if successes_in_population > population_size or sample_size > population_size:
        raise capi.DtoPanic(capi.ERR_PANIC, "Failed to create hypergeometric distribution")
    return float(eng.hypergeometric_pvalues([population_size], [successes_in_population], [sample_size], [observed_overlap])[0])


def intersect_genes(genes1, genes2) -> int:
    """src/stat_operations/intersect_genes.rs:38-56 (list-2 items whose id occurs in list 1). Host helper for
    single calls; the grid path never materialises sets -- it uses the histogram + prefix-sum kernels."""
    ids1 = {g.id() for g in genes1}
    return sum(1 for g in genes2 if g.id() in ids1)


def fdr(ranked_list_1_len: int, ranked_list_2_len: int, overlap_len: int, population_size: int, sensitivity: float) -> float:
    """src/stat_operations/fdr.rs:29-60."""
    out = C.c_double()
    capi.check(capi.lib().dto_b200_fdr(ranked_list_1_len, ranked_list_2_len, overlap_len, population_size, sensitivity, C.byref(out)))
    return out.value


def empirical_pvalue_struct(records: np.ndarray) -> capi.FinalResult:
    rec = np.ascontiguousarray(records, dtype=capi.RECORD_DTYPE)
    fin = capi.FinalResult()
    capi.check(capi.lib().dto_b200_empirical_pvalue(rec.ctypes.data_as(C.POINTER(capi.Record)), rec.size, C.byref(fin)))
    return fin


def final_json(fin: capi.FinalResult) -> str:
    n = C.c_size_t()
    L = capi.lib()
    capi.check(L.dto_b200_final_result_json(C.byref(fin), None, 0, C.byref(n)))
    buf = C.create_string_buffer(n.value + 1)
    capi.check(L.dto_b200_final_result_json(C.byref(fin), buf, n.value + 1, C.byref(n)))
    return buf.value.decode()


def empirical_pvalue(results) -> dict:
    """src/stat_operations/empirical_pvalue.rs:109-187: takes the vector of Best results (records), returns the
    JSON object (as a dict; `final_json` gives serde_json's pretty string)."""
    from .dto import records_to_array

    fin = empirical_pvalue_struct(records_to_array(results))
    return json.loads(final_json(fin))
