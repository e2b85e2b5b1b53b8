#!/usr/bin/env python
"""bench.py -- DTO permutations/sec at N = 20 000 features (BASELINE.json configs[2]) on N GPUs of one node.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...                     # the reference's CPU algorithm (oracle port) on host cores

One "step" = one pass of the hot path over one batch of synthetic input: --perms permutations (default
100 000, the configs[2] figure) PER GPU of the synthetic human-scale pair (N = 20 000, 589 x 589 thresholds),
each permutation = uniform random pairing -> overlap grid -> hypergeometric p for every threshold pair ->
minimum with the reference tie-break.  Permutation ids shard over ranks (weak scaling, no data-path collective);
the per-permutation minima are all-gathered with NCCL inside the timed region and rank 0 computes the empirical p.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs in HBM, CUDA-event timed on the
library's stream, max over ranks); `e2e` = the same metric through the public API with HOST buffers (lists
uploaded, records downloaded every step).  `roofline` = the scan kernel against the FP64 pipe under SURVEY 8(d)'s
accounting, measured live; `cpu_baseline` = the oracle on this box's host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_FEATURES = 20000
LIST_SEED = 20000
SIGMA = 0.25
PHILOX_SEED = 20000


def synthetic_lists():
    from tests import helpers as H

    return H.synthetic_pair(N_FEATURES, LIST_SEED, SIGMA)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Call at the start of the timed region: only samples taken from here on are reported."""
        self.first = len(self.rows)

    def stop(self):
        time.sleep(0.06)  # let the last 50 ms sample land
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = self.rows[self.first:] or self.rows[-1:]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_sample(faithful_stride=16, perms_per_thread=1, quick=False):
    """Times the oracle (CPU restatement of the reference path) on this box's host cores.
    faithful = string ids + per-cell HashSet + uncached Lanczos + full tails, threads like run/single_node.rs:94-133,
    on a bounded sample: one permutation per thread, every `faithful_stride`-th t1 row, extrapolated.
    optimised = integer ids, histogram + prefix sum, cached ln-factorials, converged tails (full permutations)."""
    from oracle import oracle as O
    from tests import helpers as H

    cores = os.cpu_count() or 1
    ids1, r1, ids2, r2 = synthetic_lists()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    slot = O.slot_map(o1, o2)
    n_tasks = cores * perms_per_thread
    t0 = time.perf_counter()
    O.run_single_node(o1, o2, N_FEATURES, [1] * n_tasks, cores, seed=1, mode=0, row_stride=faithful_stride)
    dt_f = time.perf_counter() - t0
    faithful = n_tasks / (dt_f * faithful_stride)
    n_opt = cores * (2 if quick else 8)
    t0 = time.perf_counter()
    O.run_single_node(o1, o2, N_FEATURES, [1] * n_opt, cores, seed=2, mode=1, slot2_of_1=slot)
    dt_o = time.perf_counter() - t0
    return {
        "value": faithful, "unit": "permutations/s", "cores": cores, "kind": "port",
        "sample": (f"reference-faithful oracle mode (string ids, per-cell hash set, uncached ln_gamma, full tails), {n_tasks} "
                   f"permutations on {cores} threads, every {faithful_stride}th t1 row of the 589x589 grid, extrapolated x{faithful_stride}; "
                   f"took {dt_f:.1f} s"),
        "optimized_port_value": n_opt / dt_o,
        "optimized_port_sample": f"integer-id oracle mode (histogram + prefix sum, cached ln-factorial), {n_opt} full permutations on {cores} threads in {dt_o:.1f} s",
    }


def algorithmic_flops_per_perm():
    """F_perm = sum over cells of (26 + 5 R), R = oracle-counted converged tail length on the same
    input (cells the reference short-circuits count 0), averaged over 2 null permutations."""
    from oracle import oracle as O
    from tests import helpers as H

    ids1, r1, ids2, r2 = synthetic_lists()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    slot = O.slot_map(o1, o2)
    lf = O.ln_factorial_table(N_FEATURES)
    tot = []
    for s in (1, 2):
        p1, p2 = H.perms(N_FEATURES, 1, s)[0], H.perms(N_FEATURES, 1, 100 + s)[0]
        g = O.grid_int(o1, o2, N_FEATURES, slot, p1, p2, lf=lf, want_p=False)
        terms, cells = O.grid_tail_terms(o1, o2, N_FEATURES, g.overlap, lf)
        tot.append(26.0 * cells + 5.0 * terms)
    return float(np.mean(tot)), int(o1.thresholds.size * o2.thresholds.size)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O  # the CPU port of the reference path (the Rust crate cannot be built here)
    from tests import helpers as H

    cores = os.cpu_count() or 1
    stride = args.ref_row_stride
    ids1, r1, ids2, r2 = synthetic_lists()
    o1, o2 = H.oracle_lists(ids1, r1, ids2, r2)
    n_tasks = cores

    def step(seed):
        O.run_single_node(o1, o2, N_FEATURES, [1] * n_tasks, cores, seed=seed, mode=0, row_stride=stride)

    for w in range(args.warmup):
        step(1000 + w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(s)
    dt = time.perf_counter() - t0
    value = n_tasks * args.steps / (dt * stride)
    unit = "permutations/s"
    line = {
        "impl": "reference", "metric": "DTO permutations/sec at N=20k features", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: synthetic human-scale lists N=20000 (589x589 threshold pairs), CPU sample",
                   "features": N_FEATURES, "threshold_pairs": int(o1.thresholds.size * o2.thresholds.size)},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": (f"reference-faithful oracle mode; each step = {n_tasks} permutations on {cores} threads, every "
                                    f"{stride}th t1 row of the grid, extrapolated x{stride} (a full N=20k permutation costs minutes of CPU)")},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--perms", type=int, default=100000, help="permutations per GPU per step (configs[2]: 100 000)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-row-stride", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # stdout must carry exactly ONE line (the JSON): native libraries write banners to fd 1 (NCCL prints its version on the
    # first communicator), so fd 1 points at stderr until the line is printed through a saved copy of the real stdout
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import __graft_entry__ as G

    if rank == 0:
        G.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    import dual_threshold_optimization_b200 as dto

    ids1, r1, ids2, r2 = synthetic_lists()
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    population = dto.compute_population_size(l1, l2, None)
    eng = dto.Engine(local_rank)  # raises without a GPU: there is no CPU path to fall back to
    if args.batch:
        eng.set_option("batch", args.batch)
    eng.load_lists(l1, l2, population)
    unperm = eng.run_unpermuted()
    T1, T2 = eng.shape[0], eng.shape[1]
    P = args.perms
    d_minp = torch.empty(P, dtype=torch.float64, device=f"cuda:{local_rank}")
    gathered = torch.empty(P * world, dtype=torch.float64, device=f"cuda:{local_rank}") if world > 1 else d_minp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(step_idx):
        first = (step_idx * world + rank) * P  # contiguous id range per rank, disjoint across steps
        eng.run_permuted_philox_device(PHILOX_SEED, first, P, d_minp.data_ptr())
        ms = eng.stats()["last_run_ms"]
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_minp)  # the small NCCL all-gather of per-permutation minima
        return ms

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # started before the warm-up (nvidia-smi needs a few 100 ms to come up); marked at the timed start
    for w in range(args.warmup):
        device_step(10_000 + w)
    barrier()
    eng.reset_stats()
    barrier()
    clocks.mark()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    scan_ms = 0.0
    sigma_ms = 0.0
    scan_launches = 0
    for s in range(args.steps):
        dev_ms += device_step(s)
        st = eng.stats()
        scan_ms += st["last_scan_kernel_ms"]
        sigma_ms += st["last_sigma_kernel_ms"]
        scan_launches += st["last_scan_launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clk = clocks.stop() if rank == 0 else None
    st = eng.stats()
    launches = st["kernel_launches"]
    emp = float((gathered <= float(unperm["pvalue"])).double().mean().item())  # host epilogue input

    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    value = P * world * args.steps / (wall_ms_max * 1e-3)  # whole job, barrier-to-barrier, max over ranks
    value_device_only = P * world * args.steps / (dev_ms_max * 1e-3)

    # ---- end-to-end through the public API with host buffers: lists up, records down, host epilogue ----
    from dual_threshold_optimization_b200._capi import RECORD_DTYPE
    from dual_threshold_optimization_b200.stat_operations import empirical_pvalue_struct

    e2e_records = np.zeros(P + 1, dtype=RECORD_DTYPE)  # host result buffer of one step: [unpermuted, P permuted]

    def e2e_step(step_idx):
        eng.load_lists(l1, l2, population)          # H2D: ranks, thresholds, slot map (screen tables: cache hit, same set sizes)
        e2e_records[0] = eng.run_unpermuted()        # D2H: the unpermuted record
        eng.run_permuted_philox(PHILOX_SEED, (step_idx * world + rank) * P, P, out=e2e_records[1:])  # D2H: P records
        return empirical_pvalue_struct(e2e_records).empirical_pvalue   # host epilogue (empirical p, FDR)

    e2e_step(20_000)
    barrier()
    eng.reset_stats()
    t0 = time.perf_counter()
    for s in range(args.e2e_steps):
        e2e_emp = e2e_step(30_000 + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    st_e = eng.stats()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = P * world * args.e2e_steps / float(t[0])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = eng.probe_fp64_tflops()
        roof = None
        cpu = None
        if not args.no_cpu_baseline:
            f_perm, cells = algorithmic_flops_per_perm()
            n1_eff = int(np.searchsorted(l1.ranks(), l1.thresholds()[-1], side='right'))
            hbm_bytes_per_perm = ((n1_eff + 1 + 255) // 256) * 256 * 2 + 40
            perms_per_launch = P * args.steps / max(scan_launches, 1)
            avg_launch_ms = scan_ms / max(scan_launches, 1)
            achieved = f_perm * perms_per_launch / (avg_launch_ms * 1e-3) / 1e12
            ncu = None
            try:
                ncu = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_ncu_latest.json")))
            except Exception:
                pass
            screened = st["level2_cells"] / max(st["tasks_fast"], 1)
            clock_hz = (clk["sm_mhz"] or 1965.0) * 1e6
            fp64_view = {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "algorithmic_flops_per_permutation": f_perm,
                "note": ("SURVEY 8(d) convention: 26 + 5R FP64 flops per evaluated cell (R = oracle-counted converged tail on the same "
                         "input). The kernel certifies-and-skips almost every cell (critical-overlap screen + log-p table), so this "
                         "algorithmic rate exceeds the pipe peak by design (SURVEY 8(d): report the post-pruning bound instead). peak = "
                         "DFMA chain measured live by dto_b200_probe_fp64_tflops (MEASURED_PEAKS.json holds no FP64 figure; nominal 37 TF)."),
            }
            hbm_view = {"bound": "hbm", "achieved": hbm_bytes_per_perm * perms_per_launch / (avg_launch_ms * 1e-3) / 1e9,
                        "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": hbm_bytes_per_perm * perms_per_launch / (avg_launch_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", float("nan")),
                        "algorithmic_bytes_per_permutation": hbm_bytes_per_perm,
                        "note": "algorithmic HBM bytes/permutation of the scan kernel = one partner-slot row read (2 B x padded n1) + one 40 B record; peak = MEASURED_PEAKS.json hbm_gbs (of measured)"}
            common = {
                "traffic": (ncu["dram_bytes_per_permutation"] * perms_per_launch) if ncu else None,
                "kernel": "dto::scan_kernel<20,true>", "avg_launch_ms": avg_launch_ms, "launches": scan_launches,
                "perms_per_launch": perms_per_launch, "share_of_step": scan_ms / dev_ms if dev_ms else None,
                "post_pruning": {
                    "cells_per_permutation": T1 * T2, "cells_past_screen_per_permutation": screened,
                    "pruning_rate": 1.0 - screened / (T1 * T2),
                    "exact_tail_evaluations_per_permutation": st["candidates"] / max(st["tasks_fast"], 1),
                    "ncu": ncu, "ncu_note": "static figures of one captured launch (profiles/scan_kernel_ncu_latest.json), not measured in this run",
                },
                "algorithmic_views": {"fp64": fp64_view, "hbm": hbm_view},
            }
            if ncu:
                # The binding resource after pruning is warp-instruction issue (FP64 pipe ~1 %, DRAM ~3 % busy in ncu): that
                # is the fraction that says how good the kernel is.  The algorithmic FP64 / HBM views SURVEY 8(d) defines
                # are kept alongside.
                inst_rate = ncu["warp_instructions_per_permutation"] * perms_per_launch / (avg_launch_ms * 1e-3)
                roof = {"bound": "issue", "achieved": inst_rate / 1e9, "peak": 148 * 4 * clock_hz / 1e9, "unit": "Gwarp-inst/s",
                        "frac": inst_rate / (148 * 4 * clock_hz),
                        "note": ("warp instructions per permutation (ncu capture) x live permutation rate of the kernel (CUDA events) vs 148 SMs x 4 "
                                 "schedulers x SM clock; neither 'hbm' nor 'tensor' binds this integer/byte path (see algorithmic_views)"),
                        **common}
            else:
                roof = {**fp64_view, **common}
            cpu = cpu_baseline_sample()
        line = {
            "metric": "DTO permutations/sec at N=20k features", "value": value, "unit": "permutations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2]: synthetic human-scale lists N=20000, 589x589 threshold pairs, on-device Philox permutations",
                       "features": N_FEATURES, "threshold_pairs": T1 * T2, "permutations_per_gpu_per_step": P,
                       "l2_policy": "inputs larger than L2: each step streams 4 GB of partner-slot rows per GPU (126 MB L2)",
                       "pvalue_evals_per_s": value * T1 * T2},
            "device_only_value": value_device_only,
            "e2e": {"value": e2e_value, "unit": "permutations/s",
                    "h2d_bytes_per_step": st_e["h2d_bytes"] // args.e2e_steps, "d2h_bytes_per_step": st_e["d2h_bytes"] // args.e2e_steps,
                    "steps": args.e2e_steps, "empirical_pvalue": e2e_emp},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_ms": {"scan": scan_ms / args.steps, "pairing_sort": sigma_ms / args.steps, "device_total": dev_ms / args.steps},
            "stats": {k: st[k] for k in ("tasks_fast", "tasks_full", "candidates", "level2_cells", "refined_cells")},
            "unpermuted": {"rank1": int(unperm["rank1"]), "rank2": int(unperm["rank2"]), "pvalue": float(unperm["pvalue"]), "empirical_pvalue": emp},
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
