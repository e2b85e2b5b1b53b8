"""Shared synthetic inputs and oracle/product glue for the tests.  Test infrastructure."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


from dual_threshold_optimization_b200.synthetic import background_subset_pair, ids_for, synthetic_pair  # noqa: E402,F401


def background_case(n_universe, n1, n2, seed):
    """two lists over different subsets of a universe, ranks with gaps/ties, background = universe."""
    rng = np.random.default_rng(seed)
    uni = ids_for(n_universe, "g")
    s1 = rng.choice(n_universe, size=n1, replace=False)
    s2 = rng.choice(n_universe, size=n2, replace=False)
    r1 = rng.integers(0, 3 * n1, size=n1).astype(np.uint32)  # rank 0 legal, gaps and ties
    r2 = rng.integers(1, 2 * n2, size=n2).astype(np.uint32)
    return [uni[i] for i in s1], r1, [uni[i] for i in s2], r2, uni


def read_csv(path):
    ids, ranks = [], []
    with open(path) as f:
        for line in f:
            a, b = line.rstrip("\n").split(",")
            ids.append(a.strip())
            ranks.append(int(b))
    return ids, np.array(ranks, dtype=np.uint32)


def load_test_data():
    ids1, r1 = read_csv(os.path.join(GOLDEN, "test_data", "ranklist1.csv"))
    ids2, r2 = read_csv(os.path.join(GOLDEN, "test_data", "ranklist2.csv"))
    bg = [ln.strip() for ln in open(os.path.join(GOLDEN, "test_data", "background.txt"))]
    return ids1, r1, ids2, r2, bg


def ten_gene_case():
    """dto/optimize_main.rs:128-157"""
    g1 = [f"gene{i}" for i in range(1, 11)]
    g2 = ["gene7", "gene3", "gene9", "gene1", "gene5", "gene10", "gene4", "gene8", "gene6", "gene2"]
    r = np.arange(1, 11, dtype=np.uint32)
    return g1, r, g2, r.copy()


def oracle_lists(ids1, r1, ids2, r2):
    return O.OracleRankedList.make(ids1, r1), O.OracleRankedList.make(ids2, r2)


def perms(n, P, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(n).astype(np.uint32) for _ in range(P)])


def perms_from_pairing(pairing, slot2_of_1, n2):
    """Builds (perm1, perm2) index vectors (permuted.rs semantics) that realise a device pairing
    pos2_of_pos1[j] (0xFFFFFFFF = position j holds a gene absent from list 2)."""
    n1 = pairing.size
    slot2_of_1 = np.asarray(slot2_of_1)
    common1 = [a for a in range(n1) if slot2_of_1[a] >= 0]
    other1 = [a for a in range(n1) if slot2_of_1[a] < 0]
    paired_pos = [j for j in range(n1) if pairing[j] != 0xFFFFFFFF]
    assert len(paired_pos) == len(common1)
    perm1 = np.empty(n1, dtype=np.uint32)
    perm2 = np.full(n2, 0xFFFFFFFF, dtype=np.uint32)
    for t, j in enumerate(paired_pos):
        a = common1[t]
        perm1[j] = a
        perm2[pairing[j]] = slot2_of_1[a]
    unp = [j for j in range(n1) if pairing[j] == 0xFFFFFFFF]
    for j, a in zip(unp, other1):
        perm1[j] = a
    used2 = set(int(x) for x in perm2 if x != 0xFFFFFFFF)
    rest2 = [b for b in range(n2) if b not in used2]
    free2 = [j for j in range(n2) if perm2[j] == 0xFFFFFFFF]
    for j, b in zip(free2, rest2):
        perm2[j] = b
    return perm1, perm2


def assert_record_matches(rec, ob, rel=1e-12):
    """integer fields bit-exact; p within `rel` relative (exact when 0)."""
    for f in ("rank1", "rank2", "set1_len", "set2_len", "intersection_size", "population_size"):
        assert int(rec[f]) == int(ob[f]), (f, rec, ob)
    p, q = float(rec["pvalue"]), float(ob["pvalue"])
    if q == 0.0:
        assert p == 0.0, (p, q)
    else:
        assert abs(p - q) <= rel * abs(q), (p, q)


# ---------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. SC'11) in numpy: an independent restatement used to replay the device generator
# ---------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & MASK, p0 & MASK
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def device_keys(n, seed, perm_id, stream):
    """One 32-bit Philox word per element: counter = (element / 4, stream, id lo, id hi), key = (seed lo, seed hi),
    word element % 4 (the SECONDARY sort key of the pairing kernel uses stream + 8)."""
    blk = np.arange((n + 3) // 4, dtype=np.uint64)
    out = philox4x32_10(blk, np.full_like(blk, stream), np.full_like(blk, perm_id & 0xFFFFFFFF), np.full_like(blk, perm_id >> 32),
                        seed & 0xFFFFFFFF, seed >> 32)
    return np.stack(out, axis=1).reshape(-1)[:n]


def device_keys16(n, seed, perm_id, stream):
    """The 16-bit PRIMARY sort key of every element: one Philox call per 8 elements (counter = (element / 8, stream, id lo,
    id hi)); element e takes word (e % 8) / 2, low half when e is even, high half when odd."""
    blk = np.arange((n + 7) // 8, dtype=np.uint64)
    out = philox4x32_10(blk, np.full_like(blk, stream), np.full_like(blk, perm_id & 0xFFFFFFFF), np.full_like(blk, perm_id >> 32),
                        seed & 0xFFFFFFFF, seed >> 32)
    words = np.stack(out, axis=1)  # [n8, 4]
    halves = np.stack([words & np.uint64(0xFFFF), words >> np.uint64(16)], axis=2)  # [n8, 4, 2]
    return halves.reshape(-1)[:n]


def expected_pairing_identical(n, seed, perm_id):
    """pos2_of_pos1 for identical gene sets: list-1 position f is paired with the list-2 position whose key has rank f.
    Order = (16-bit primary key, top 16 + B bits of the 32-bit secondary Philox key (stream + 8), index), B = bucket bits
    of the kernel (the tie-break word packs the key16 bits below the bucket bits and those secondary bits into 32 bits)."""
    lg = 0
    while (1 << lg) < n:
        lg += 1
    B = min(14, max(1, lg - 1))
    key = device_keys16(n, seed, perm_id, 0)
    sec = device_keys(n, seed, perm_id, 8) >> np.uint64(16 - B)
    order = np.lexsort((np.arange(n), sec, key))
    return order.astype(np.uint32)
