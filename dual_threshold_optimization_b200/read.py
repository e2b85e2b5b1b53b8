"""Mirror of the reference's `read` module (src/read/*.rs)."""
from __future__ import annotations

import ctypes as C

from . import _capi as capi
from .collections import FeatureList, RankedFeatureList


def read_ranked_feature_list_from_csv(filepath: str) -> RankedFeatureList:
    """src/read/read_ranked_feature_list_from_csv.rs:50-70: `feature,rank` per line, no header, exactly two
    fields, fields trimmed, rank parsed as usize then truncated to u32.  The reference's panics raise DtoPanic."""
    h = C.c_void_p()
    capi.check(capi.lib().dto_b200_read_ranked_list_csv(str(filepath).encode(), C.byref(h)))
    return RankedFeatureList(h)


def read_feature_list_from_file(filepath: str) -> FeatureList:
    """src/read/read_feature_list_from_file.rs:45-53: one trimmed id per line (blank lines count)."""
    h = C.c_void_p()
    L = capi.lib()
    capi.check(L.dto_b200_read_feature_list(str(filepath).encode(), C.byref(h)))
    try:
        n = L.dto_b200_feature_list_len(h)
        ids = [L.dto_b200_feature_list_id(h, i).decode() for i in range(n)]  # the C reader is the only parser of the format
    finally:
        L.dto_b200_feature_list_free(h)
    return FeatureList(ids)
