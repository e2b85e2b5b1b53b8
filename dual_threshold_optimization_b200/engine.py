"""Low-level engine: one context per GPU over the integer-id C ABI (include/dto_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi as capi


class Engine:
    """Owns a dto_b200_ctx bound to one CUDA device.  Fails loudly (DtoError) without a usable GPU."""

    def __init__(self, device: int = 0):
        self._ctx = C.c_void_p()
        capi.check(capi.lib().dto_b200_create(C.byref(self._ctx), device))
        self.device = device
        self.shape = None  # (T1, T2, n1, n2)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            capi.lib().dto_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def ctx(self):
        return self._ctx

    def set_option(self, name: str, value: int):
        capi.check(capi.lib().dto_b200_set_option(self._ctx, name.encode(), int(value)))

    def set_problem(self, ranks1, thr1, ranks2, thr2, slot2_of_1, population: int):
        r1, t1, r2, t2 = capi.u32(ranks1), capi.u32(thr1), capi.u32(ranks2), capi.u32(thr2)
        sm = np.ascontiguousarray(slot2_of_1, dtype=np.int32)
        if sm.size != r1.size:
            raise ValueError("slot2_of_1 must have one entry per list-1 feature")
        capi.check(
            capi.lib().dto_b200_set_problem(
                self._ctx, capi.ptr(r1, C.c_uint32), r1.size, capi.ptr(t1, C.c_uint32), t1.size,
                capi.ptr(r2, C.c_uint32), r2.size, capi.ptr(t2, C.c_uint32), t2.size,
                capi.ptr(sm, C.c_int32), int(population),
            )
        )
        self.shape = (t1.size, t2.size, r1.size, r2.size)

    def load_lists(self, l1, l2, population: int):
        capi.check(capi.lib().dto_b200_load_lists(self._ctx, l1.handle, l2.handle, int(population)))
        self.shape = (len(l1.thresholds()), len(l2.thresholds()), len(l1), len(l2))

    def run_unpermuted(self) -> np.void:
        rec = np.zeros(1, dtype=capi.RECORD_DTYPE)
        capi.check(capi.lib().dto_b200_run_unpermuted(self._ctx, rec.ctypes.data_as(C.POINTER(capi.Record))))
        return rec[0]

    def run_permuted_indices(self, perm1, perm2) -> np.ndarray:
        p1, p2 = capi.u32(perm1), capi.u32(perm2)
        if p1.ndim != 2 or p2.ndim != 2 or p1.shape[0] != p2.shape[0]:
            raise ValueError("perm1/perm2 must be P x n1 and P x n2")
        if self.shape is None or p1.shape[1] != self.shape[2] or p2.shape[1] != self.shape[3]:
            raise ValueError("perm row length does not match the loaded lists")
        P = p1.shape[0]
        rec = np.zeros(P, dtype=capi.RECORD_DTYPE)
        capi.check(
            capi.lib().dto_b200_run_permuted_indices(
                self._ctx, capi.ptr(p1, C.c_uint32), capi.ptr(p2, C.c_uint32), P, rec.ctypes.data_as(C.POINTER(capi.Record))
            )
        )
        return rec

    def run_permuted_philox(self, seed: int, first_perm_id: int, P: int, want_records=True, want_minp=False, out=None):
        """P device-generated permutations.  `out` (optional): a contiguous RECORD_DTYPE array of >= P records to fill in
        place (e.g. a slice of a larger buffer), instead of a fresh allocation."""
        if out is not None:
            if out.dtype != capi.RECORD_DTYPE or out.size < P or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a contiguous RECORD_DTYPE array with at least P records")
            rec = out
        else:
            rec = np.zeros(P, dtype=capi.RECORD_DTYPE) if want_records else None
        mp = np.zeros(P, dtype=np.float64) if want_minp else None
        capi.check(
            capi.lib().dto_b200_run_permuted_philox(
                self._ctx, int(seed), int(first_perm_id), int(P),
                rec.ctypes.data_as(C.POINTER(capi.Record)) if rec is not None else None,
                capi.ptr(mp, C.c_double),
            )
        )
        if rec is not None and want_minp:
            return rec, mp
        return rec if rec is not None else mp

    def run_permuted_philox_device(self, seed: int, first_perm_id: int, P: int, d_minp_ptr: int, d_records_ptr: int = 0):
        """Outputs stay on the device (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        capi.check(
            capi.lib().dto_b200_run_permuted_philox_device(
                self._ctx, int(seed), int(first_perm_id), int(P), C.c_void_p(d_minp_ptr or None), C.c_void_p(d_records_ptr or None)
            )
        )

    def philox_pairing(self, seed: int, perm_id: int) -> np.ndarray:
        out = np.zeros(self.shape[2], dtype=np.uint32)
        capi.check(capi.lib().dto_b200_philox_pairing(self._ctx, int(seed), int(perm_id), capi.ptr(out, C.c_uint32)))
        return out

    def grid_debug(self, perm1=None, perm2=None, want_p=True, want_logp=False):
        T1, T2 = self.shape[0], self.shape[1]
        ov = np.zeros((T1, T2), dtype=np.uint32)
        pv = np.zeros((T1, T2), dtype=np.float64) if want_p else None
        lp = np.zeros((T1, T2), dtype=np.float64) if want_logp else None
        p1 = None if perm1 is None else capi.u32(perm1)
        p2 = None if perm2 is None else capi.u32(perm2)
        capi.check(
            capi.lib().dto_b200_grid_debug(
                self._ctx, capi.ptr(p1, C.c_uint32), capi.ptr(p2, C.c_uint32), capi.ptr(ov, C.c_uint32),
                capi.ptr(pv, C.c_double), capi.ptr(lp, C.c_double),
            )
        )
        return ov, pv, lp

    def hypergeometric_pvalues(self, N, K, n, k) -> np.ndarray:
        arrs = [np.ascontiguousarray(a, dtype=np.uint64).ravel() for a in (N, K, n, k)]
        cnt = arrs[0].size
        if any(a.size != cnt for a in arrs):
            raise ValueError("N, K, n, k must have equal lengths")
        out = np.zeros(cnt, dtype=np.float64)
        capi.check(
            capi.lib().dto_b200_hypergeometric_pvalues(
                self._ctx, *[capi.ptr(a, C.c_uint64) for a in arrs], cnt, capi.ptr(out, C.c_double)
            )
        )
        return out

    def stats(self) -> dict:
        s = capi.Stats()
        capi.check(capi.lib().dto_b200_get_stats(self._ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def last_batch_task_stats(self, max_tasks: int = 1 << 20) -> np.ndarray:
        """(n, 8) uint32 per task of the last launch: screened, refined, exact cells; cycles/16 of task, scatter, drain,
        refine, exact stages."""
        out = np.zeros((max_tasks, 8), dtype=np.uint32)
        n = C.c_size_t()
        capi.check(capi.lib().dto_b200_last_batch_task_stats(self._ctx, capi.ptr(out, C.c_uint32), max_tasks, C.byref(n)))
        return out[: n.value]

    def table_logp(self, rows, cols, ks) -> np.ndarray:
        """Entries of the device log-p lookup table (NaN outside a cell's tabulated range) -- diagnostics."""
        r, c, k = capi.u32(rows), capi.u32(cols), capi.u32(ks)
        out = np.zeros(r.size, dtype=np.float64)
        capi.check(capi.lib().dto_b200_table_logp(self._ctx, capi.ptr(r, C.c_uint32), capi.ptr(c, C.c_uint32),
                                                  capi.ptr(k, C.c_uint32), r.size, capi.ptr(out, C.c_double)))
        return out

    def reset_stats(self):
        capi.check(capi.lib().dto_b200_reset_stats(self._ctx))

    def probe_fp64_tflops(self) -> float:
        v = C.c_double()
        capi.check(capi.lib().dto_b200_probe_fp64_tflops(self._ctx, C.byref(v)))
        return v.value

    def probe_hbm_gbs(self) -> float:
        v = C.c_double()
        capi.check(capi.lib().dto_b200_probe_hbm_gbs(self._ctx, C.byref(v)))
        return v.value


def device_count() -> int:
    n = C.c_int()
    capi.check(capi.lib().dto_b200_device_count(C.byref(n)))
    return n.value
