"""Permutation-id sharding across GPUs and the gather of per-permutation minima.

Replaces the MPI scatter/gather of src/run/multi_node.rs:114-161 on one box: tasks are independent, so rank g
owns the contiguous permutation-id range [g*P/G, (g+1)*P/G) (ids, not data, are scattered -- every rank holds its
own copy of the small problem tables), and the only exchange is one all-gather of P doubles (NCCL over NVLink on
GPUs, gloo in the CPU tests).  The Philox pairing is a pure function of (seed, id), so results do not depend on G.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_perm: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, first+count) of permutation ids owned by `rank`: ceil-sized contiguous chunks like
    tasks.chunks(len.div_ceil(ranks)) in the reference (multi_node.rs:117-120)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-n_perm // world) if n_perm else 0
    first = min(rank * per, n_perm)
    return first, min(per, n_perm - first)


def gather_minima(local_minp, n_perm: int, group=None):
    """All-gathers the per-rank minima (torch tensor of this rank's shard, any device) into the full id-ordered
    vector of length n_perm on every rank.  Shards are padded to the ceil size so one all_gather_into_tensor
    suffices; padding is +inf and is stripped."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_minp[:n_perm]
    per = -(-n_perm // world)
    pad = torch.full((per,), float("inf"), dtype=local_minp.dtype, device=local_minp.device)
    pad[: local_minp.numel()] = local_minp
    out = torch.empty(per * world, dtype=local_minp.dtype, device=local_minp.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        first, cnt = shard_range(n_perm, world, r)
        pieces.append(out[r * per : r * per + cnt])
    return torch.cat(pieces)


def empirical_from_minima(minima, unpermuted_p: float) -> float:
    """stat_operations/empirical_pvalue.rs:160-165: #{p_perm <= p_unperm} / P, no +1; P == 0 -> 1.0."""
    m = np.asarray(minima, dtype=np.float64)
    if m.size == 0:
        return 1.0
    return float(np.count_nonzero(m <= unpermuted_p)) / float(m.size)
