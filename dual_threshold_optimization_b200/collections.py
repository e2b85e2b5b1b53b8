"""Mirror of the reference's `collections` module (src/collections/*.rs) over the C-ABI host layer."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence

import numpy as np

from . import _capi as capi


class Feature:
    """src/collections/feature.rs:25-77 -- a feature is its string id."""

    __slots__ = ("_id",)

    def __init__(self, id: str):
        self._id = str(id)

    @classmethod
    def from_(cls, id: str) -> "Feature":
        return cls(id)

    def id(self) -> str:
        return self._id

    def __eq__(self, o):
        return isinstance(o, Feature) and o._id == self._id

    def __hash__(self):
        return hash(self._id)

    def __repr__(self):
        return f"Feature({self._id!r})"


class FeatureList:
    """src/collections/feature_list.rs -- ordered list of features (background lists)."""

    def __init__(self, genes: Iterable = ()):
        self._ids: List[str] = [g.id() if isinstance(g, Feature) else str(g) for g in genes]
        self._handle = None

    @classmethod
    def from_(cls, genes: Iterable) -> "FeatureList":
        return cls(genes)

    def genes(self) -> List[Feature]:
        return [Feature(g) for g in self._ids]

    def ids(self) -> List[str]:
        return list(self._ids)

    def __len__(self):
        return len(self._ids)

    def __iter__(self):
        return iter(self.genes())

    @property
    def handle(self):
        if self._handle is None:
            h = C.c_void_p()
            capi.check(capi.lib().dto_b200_feature_list_from(capi.c_strings(self._ids), len(self._ids), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None and self._handle.value:
                capi.lib().dto_b200_feature_list_free(self._handle)
        except Exception:
            pass


class RankedFeatureList:
    """src/collections/ranked.rs:125-134.  Construction = RankedFeatureList::from (:176-191): length check,
    stable sort by rank, threshold series 1, floor(t*1.01+1), ... while <= max rank (last threshold NOT forced
    to the max rank -- the reference's :370-372 statement is a no-op and is reproduced)."""

    def __init__(self, handle: C.c_void_p):
        self._handle = handle

    @classmethod
    def from_(cls, genes, ranks: Sequence[int]) -> "RankedFeatureList":
        ids = [g.id() if isinstance(g, Feature) else str(g) for g in (genes.genes() if isinstance(genes, FeatureList) else genes)]
        r = capi.u32(ranks)
        if len(ids) != r.size:
            # check_lengths, ranked.rs:515-525
            raise ValueError(f"Genes and ranks must have the same length. Genes: {len(ids)}, Ranks: {r.size}")
        h = C.c_void_p()
        capi.check(capi.lib().dto_b200_ranked_list_from(capi.c_strings(ids), capi.ptr(r, C.c_uint32), len(ids), C.byref(h)))
        return cls(h)

    @property
    def handle(self):
        return self._handle

    def __len__(self):
        return capi.lib().dto_b200_ranked_list_len(self._handle)

    def len(self):
        return len(self)

    def is_empty(self):
        return len(self) == 0

    def ranks(self) -> np.ndarray:
        n = len(self)
        p = capi.lib().dto_b200_ranked_list_ranks(self._handle)
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)

    def thresholds(self) -> np.ndarray:
        n = capi.lib().dto_b200_ranked_list_num_thresholds(self._handle)
        p = capi.lib().dto_b200_ranked_list_thresholds(self._handle)
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)

    def ids(self) -> List[str]:
        L = capi.lib()
        return [L.dto_b200_ranked_list_id(self._handle, i).decode() for i in range(len(self))]

    def genes(self) -> FeatureList:
        return FeatureList(self.ids())

    def get_feature_set_by_threshold(self, threshold: int) -> List[Feature]:
        """FeatureSetProvider (ranked.rs:561-568): genes whose rank <= threshold."""
        r = self.ranks()
        ids = self.ids()
        return [Feature(ids[j]) for j in range(len(ids)) if r[j] <= threshold]

    def __del__(self):
        try:
            if self._handle is not None and self._handle.value:
                capi.lib().dto_b200_ranked_list_free(self._handle)
        except Exception:
            pass


class PermutedRankedFeatureList:
    """src/collections/permuted.rs:30-101 -- a permuted *view*: sorted position j keeps ranks[j] and holds the
    gene of slot indices[j].  The reference draws `indices` from thread_rng (unseedable); here they are either
    supplied by the caller (parity mode) or drawn with numpy's PCG64 from `seed`."""

    def __init__(self, original: RankedFeatureList, indices=None, seed=None):
        self.original = original
        n = len(original)
        if indices is None:
            indices = np.random.default_rng(seed).permutation(n)
        self.indices = capi.u32(indices)
        if self.indices.size != n or not np.array_equal(np.sort(self.indices), np.arange(n, dtype=np.uint32)):
            raise ValueError("indices must be a permutation of 0..len-1")

    def thresholds(self):
        return self.original.thresholds()

    def ranks(self):
        return self.original.ranks()

    def get_feature_set_by_threshold(self, threshold: int) -> List[Feature]:
        r = self.original.ranks()
        ids = self.original.ids()
        return [Feature(ids[self.indices[j]]) for j in range(len(ids)) if r[j] <= threshold]
