"""Synthetic ranked lists of the BASELINE.json shapes: numpy only, shared by bench.py, the tools and the
tests.  Lives in the package (not under tests/) so that a product process never has to import test infrastructure."""
from __future__ import annotations

import numpy as np


def ids_for(n, prefix="f"):
    return [f"{prefix}{i + 1:06d}" for i in range(n)]


def synthetic_pair(N, seed, sigma=0.25, tied_frac=0.0):
    """list 1: rank(i) = i+1; list 2: rank of i + Normal(0, sigma*N) (sigma=None -> pure shuffle, 0 -> identical)."""
    rng = np.random.default_rng(seed)
    ids = ids_for(N)
    r1 = np.arange(1, N + 1, dtype=np.uint32)
    if sigma is None:
        score = rng.permutation(N).astype(np.float64)
    else:
        score = np.arange(N) + rng.normal(0.0, sigma * N if sigma > 0 else 0.0, N)
    order = np.argsort(score, kind="stable")
    r2 = np.empty(N, dtype=np.uint32)
    r2[order] = np.arange(1, N + 1, dtype=np.uint32)
    if tied_frac > 0:
        # collapse a fraction of features into tie groups with `min` ranking
        for r in (r1, r2):
            n_groups = max(1, int(N * tied_frac / 4))
            starts = rng.choice(np.arange(1, N - 4), size=n_groups, replace=False)
            for s in starts:
                sel = (r >= s) & (r < s + 4)
                r[sel] = s
    return ids, r1, list(ids), r2


def background_subset_pair(universe=60000, background=40000, seed=60000, sigma=0.3):
    """BASELINE configs[4]: both lists rank a `universe` of ids, then are filtered to a random `background` subset keeping
    their ORIGINAL ranks (gaps), population = |background|.  Returns (ids1, ranks1, ids2, ranks2, background_ids)."""
    rng = np.random.default_rng(seed)
    ids1, r1, ids2, r2 = synthetic_pair(universe, seed, sigma)
    keep = np.zeros(universe, dtype=bool)
    keep[rng.choice(universe, size=background, replace=False)] = True
    # ids are f000001.. in universe order for BOTH lists (list 2 differs in ranks only)
    sel = np.nonzero(keep)[0]
    f1 = [ids1[i] for i in sel]
    f2 = [ids2[i] for i in sel]
    return f1, r1[sel], f2, r2[sel], [ids1[i] for i in sel]


def pair_batch(n_pairs, N=6000, first_seed=1, sigma_cycle=(0.25, None, 0.35, None)):
    """BASELINE configs[3]: independent list pairs over the same N features (seeds first_seed ..), a mix of concordant
    and null pairs.  Yields (ids1, ranks1, ids2, ranks2)."""
    for q in range(n_pairs):
        yield synthetic_pair(N, first_seed + q, sigma_cycle[q % len(sigma_cycle)])
