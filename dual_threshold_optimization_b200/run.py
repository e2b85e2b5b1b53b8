"""Mirror of the reference's `run` module (src/run/*.rs)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _capi as capi
from .collections import RankedFeatureList
from .dto import OptimizationResultRecord


@dataclass
class Task:
    """src/run/task.rs:3-9"""

    id: int
    permute: bool


def run_single_node_records(tasks: Sequence[Task], l1: RankedFeatureList, l2: RankedFeatureList, population_size: int,
                            num_threads: int = 1, devices: Optional[Sequence[int]] = None, seed: int = 0,
                            out: Optional[np.ndarray] = None) -> np.ndarray:
    if isinstance(tasks, tuple) and len(tasks) == 2 and isinstance(tasks[0], np.ndarray):
        ids, tp = np.ascontiguousarray(tasks[0], dtype=np.uint64), np.ascontiguousarray(tasks[1], dtype=np.uint8)  # (ids, permute) arrays
    else:
        ids = np.array([t.id for t in tasks], dtype=np.uint64)
        tp = np.array([1 if t.permute else 0 for t in tasks], dtype=np.uint8)
    if out is None:
        out = np.zeros(tp.size, dtype=capi.RECORD_DTYPE)
    elif out.dtype != capi.RECORD_DTYPE or out.size < tp.size or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be a contiguous RECORD_DTYPE array with one record per task")
    devs = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
    capi.check(
        capi.lib().dto_b200_run_tasks(
            l1.handle, l2.handle, int(population_size), ids.ctypes.data_as(C.POINTER(C.c_uint64)),
            tp.ctypes.data_as(C.POINTER(C.c_uint8)), tp.size,
            devs.ctypes.data_as(C.POINTER(C.c_int)) if devs.size else None, devs.size, int(seed),
            out.ctypes.data_as(C.POINTER(capi.Record)),
        )
    )
    return out[: tp.size]


def run_single_node(tasks: Sequence[Task], l1: RankedFeatureList, l2: RankedFeatureList, population_size: int,
                    num_threads: int = 1, devices: Optional[Sequence[int]] = None, seed: int = 0) -> List[OptimizationResultRecord]:
    """src/run/single_node.rs:83-137.  `num_threads` is accepted for signature compatibility; tasks are batched
    onto the GPU(s) in `devices` (default: device 0).  Results come back in task order (the reference returns them
    in thread-completion order; its only consumer partitions by `permuted`).  A permuted task draws the device
    permutation with Philox id Task.id under `seed`, so a job sharded over processes (each with its slice of the tasks)
    yields the records of the unsharded run."""
    return [OptimizationResultRecord.from_np(r) for r in run_single_node_records(tasks, l1, l2, population_size, num_threads, devices, seed)]


def run_multi_gpu(tasks, l1, l2, population_size, devices=None, seed: int = 0):
    """Single-box replacement of run_multi_node (src/run/multi_node.rs:47-162): shard over all visible GPUs."""
    from .engine import device_count

    devs = list(devices) if devices is not None else list(range(device_count()))
    return run_single_node(tasks, l1, l2, population_size, 1, devs, seed)


def run_pairs(pairs, permutations: int, devices: Optional[Sequence[int]] = None, seed: int = 0) -> List[dict]:
    """Batched list-pair driver (BASELINE config 4): `pairs` is a sequence of (list1, list2, population_size); each pair
    gets the whole CLI run of src/main.rs:91-165 (unpermuted optimum, `permutations` permuted tasks, empirical p, FDR)
    and yields the reference's JSON object as a dict.  Pairs shard over `devices`."""
    import json

    from .stat_operations import final_json

    n = len(pairs)
    L1 = (C.c_void_p * max(n, 1))(*[p[0].handle for p in pairs])
    L2 = (C.c_void_p * max(n, 1))(*[p[1].handle for p in pairs])
    pops = np.ascontiguousarray([int(p[2]) for p in pairs], dtype=np.uint64)
    out = (capi.FinalResult * max(n, 1))()
    devs = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
    capi.check(
        capi.lib().dto_b200_run_pairs(
            L1, L2, pops.ctypes.data_as(C.POINTER(C.c_uint64)), n, int(permutations),
            devs.ctypes.data_as(C.POINTER(C.c_int)) if devs.size else None, devs.size, int(seed), out,
        )
    )
    return [json.loads(final_json(out[i])) for i in range(n)]


def run_pairs_structs(pairs, permutations: int, devices: Optional[Sequence[int]] = None, seed: int = 0):
    """run_pairs without the JSON round trip: the ctypes array of dto_b200_final_result structs."""
    n = len(pairs)
    L1 = (C.c_void_p * max(n, 1))(*[p[0].handle for p in pairs])
    L2 = (C.c_void_p * max(n, 1))(*[p[1].handle for p in pairs])
    pops = np.ascontiguousarray([int(p[2]) for p in pairs], dtype=np.uint64)
    out = (capi.FinalResult * max(n, 1))()
    devs = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
    capi.check(
        capi.lib().dto_b200_run_pairs(
            L1, L2, pops.ctypes.data_as(C.POINTER(C.c_uint64)), n, int(permutations),
            devs.ctypes.data_as(C.POINTER(C.c_int)) if devs.size else None, devs.size, int(seed), out,
        )
    )
    return out
