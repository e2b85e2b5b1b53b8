#!/usr/bin/env python
"""bench.py -- DTO permutations/sec (BASELINE.json) on N GPUs of one node.

  python bench.py [--gpus 1] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--scaling weak|strong]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W [...]
  python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on the host cores

Workloads (BASELINE.json configs; the default, c3, is the one the headline metric is quoted on):
  c2  N = 6 000 (469 x 469 threshold pairs),  10 000 permutations per step
  c3  N = 20 000 (589 x 589),                100 000 permutations per step
  c4  2 000 list pairs of N = 6 000, 1 000 permutations (+ the unpermuted task) per pair, sharded by pair
  c5  60 000-id universe filtered to a 40 000-id background (700 x 700), 1 000 000 permutations over 8 GPUs
One "step" = one pass of the hot path over one batch: every permutation = uniform random pairing -> overlap grid ->
hypergeometric p for every threshold pair -> minimum with the reference tie-break.  --scaling weak (default): the
per-step batch above is PER GPU (c5: 125 000 per GPU); strong: it is the whole job, permutation ids sharded over ranks.
c4 is always "strong" (the batch of pairs is the job).

Prints ONE JSON line (rank 0).  `value` = throughput with the lists resident in HBM (device-timed call through the
engine layer of the C ABI; the per-permutation minima are all-gathered over NCCL inside the timed region when N > 1);
`e2e` = the same metric through the drop-in boundary with HOST buffers: dto_b200_run_tasks / dto_b200_run_pairs on list
handles (string ids canonicalised, lists uploaded, records downloaded) + dto_b200_empirical_pvalue, every step.
`roofline`, `cpu_baseline`: see DESIGN.md section 6.
"""
import argparse
import hashlib
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

from dual_threshold_optimization_b200 import synthetic as S  # noqa: E402  (numpy only)

PHILOX_SEED = 20000
METRIC = "DTO permutations/sec at N=20k features"

CONFIGS = {
    "c2": {"label": "configs[1]: synthetic yeast-scale lists N=6000 (469x469 threshold pairs), 10000 permutations", "features": 6000,
           "list_seed": 6000, "sigma": 0.25, "perms": 10000},
    "c3": {"label": "configs[2]: synthetic human-scale lists N=20000 (589x589 threshold pairs), 100000 permutations", "features": 20000,
           "list_seed": 20000, "sigma": 0.25, "perms": 100000},
    "c4": {"label": "configs[3]: batch of 2000 list pairs (N=6000 each), 1000 permutations + the unpermuted task per pair", "features": 6000,
           "pairs": 2000, "perms_per_pair": 1000},
    "c5": {"label": "configs[4]: 60000-id universe filtered to a 40000-id background (700x700 threshold pairs), 1000000 permutations over 8 GPUs",
           "universe": 60000, "background": 40000, "list_seed": 60000, "sigma": 0.3, "perms": 125000, "perms_strong": 1000000},
}


def workload_lists(cfg_name):
    """(ids1, ranks1, ids2, ranks2, background or None) of the single-pair workloads."""
    c = CONFIGS[cfg_name]
    if cfg_name == "c5":
        return S.background_subset_pair(c["universe"], c["background"], c["list_seed"], c["sigma"])
    ids1, r1, ids2, r2 = S.synthetic_pair(c["features"], c["list_seed"], c["sigma"])
    return ids1, r1, ids2, r2, None


def kernel_source_sha():
    """sha256 over the kernel sources: an ncu capture under profiles/ only describes the build it was taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "dual_threshold_optimization_b200", "csrc")
    for f in ("dto_kernels.cu", "dto_kernels.cuh", "dto_device.cuh"):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the oracle (CPU restatement of the reference path).  Only the cpu_baseline leg and the reference arm execute
# it, and they do so in a CHILD process, so the process that times the CUDA path never maps the oracle library.
# ---------------------------------------------------------------------------------------------------------------------
def cpu_worker(args):
    """Child process: prints one JSON object.  mode 'baseline': bounded CPU sample + algorithmic flops of the workload;
    mode 'reference': the timed loop of the reference arm."""
    from oracle import oracle as O

    cfg = args.config
    feats = CONFIGS[cfg]["features"] if cfg != "c5" else CONFIGS[cfg]["background"]
    if cfg == "c4":
        ids1, r1, ids2, r2 = S.synthetic_pair(feats, 1, 0.25)
        bg = None
    else:
        ids1, r1, ids2, r2, bg = workload_lists(cfg)
    o1, o2 = O.OracleRankedList.make(ids1, r1), O.OracleRankedList.make(ids2, r2)
    pop = O.compute_population_size(o1, o2, bg)
    cores = os.cpu_count() or 1
    stride = args.ref_row_stride
    out = {"cores": cores, "features": feats, "threshold_pairs": int(o1.thresholds.size * o2.thresholds.size)}
    if args.cpu_worker == "reference":
        def step(seed):
            O.run_single_node(o1, o2, pop, [1] * cores, cores, seed=seed, mode=0, row_stride=stride)

        for w in range(args.warmup):
            step(1000 + w)
        t0 = time.perf_counter()
        for s in range(args.steps):
            step(s)
        dt = time.perf_counter() - t0
        out.update({"value": cores * args.steps / (dt * stride), "seconds": dt, "perms_per_step": cores, "stride": stride})
    else:
        slot = O.slot_map(o1, o2)
        t0 = time.perf_counter()
        O.run_single_node(o1, o2, pop, [1] * cores, cores, seed=1, mode=0, row_stride=stride)
        dt_f = time.perf_counter() - t0
        n_opt = cores * (2 if feats > 30000 else 6)
        t0 = time.perf_counter()
        O.run_single_node(o1, o2, pop, [1] * n_opt, cores, seed=2, mode=1, slot2_of_1=slot)
        dt_o = time.perf_counter() - t0
        # F_perm = sum over cells of (26 + 5 R), R = oracle-counted converged tail length on the same input
        lf = O.ln_factorial_table(pop)
        rng = np.random.default_rng(1)
        tot = []
        for _ in range(2):
            p1, p2 = rng.permutation(len(ids1)).astype(np.uint32), rng.permutation(len(ids2)).astype(np.uint32)
            g = O.grid_int(o1, o2, pop, slot, p1, p2, lf=lf, want_p=False)
            terms, cells = O.grid_tail_terms(o1, o2, pop, g.overlap, lf)
            tot.append(26.0 * cells + 5.0 * terms)
        out.update({"faithful_value": cores / (dt_f * stride), "faithful_seconds": dt_f, "stride": stride,
                    "optimized_value": n_opt / dt_o, "optimized_seconds": dt_o, "optimized_perms": n_opt,
                    "algorithmic_flops_per_permutation": float(np.mean(tot))})
    print(json.dumps(out), flush=True)


def run_cpu_worker(mode, args):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", mode, "--config", args.config, "--steps", str(args.steps),
           "--warmup", str(args.warmup), "--ref-row-stride", str(args.ref_row_stride)]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    if r.returncode != 0:
        raise RuntimeError(f"cpu worker failed: {r.stderr[-2000:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def workload_string(args, cfg_name):
    return CONFIGS[cfg_name]["label"]


def nominal_permutations_per_step(args):
    """permutations one step of the workload stands for, over all GPUs (what `config.permutations_per_step` says)"""
    c = CONFIGS[args.config]
    if args.config == "c4":
        return (args.pairs or c["pairs"]) * (args.perms or c["perms_per_pair"])
    if args.scaling == "strong":
        return args.perms or c.get("perms_strong", c["perms"])
    return (args.perms or c["perms"]) * max(args.gpus, 1)


def run_reference_arm(args, rank):
    """The reference's own CPU algorithm for this path on the box's host cores: the oracle port in reference-faithful mode
    (the Rust crate cannot be built in this image), every host thread, a bounded sample per step."""
    if rank != 0:
        return
    w = run_cpu_worker("reference", args)
    value = w["value"]
    unit = "permutations/s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * w["seconds"] / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling if args.config != "c4" else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        # the workload of the product arm (same keys, same values); what one step of THIS arm actually times -- a bounded
        # sample of it -- is described in cpu_baseline.sample
        "config": {"workload": workload_string(args, args.config), "features": w["features"], "threshold_pairs": w["threshold_pairs"],
                   "permutations_per_step": nominal_permutations_per_step(args)},
        "cpu_baseline": {"value": value, "unit": unit, "cores": w["cores"], "kind": "port",
                         "sample": (f"reference-faithful oracle mode (string ids, per-cell hash set, uncached ln_gamma, full tails); each step = "
                                    f"{w['perms_per_step']} permutations on {w['cores']} threads, every {w['stride']}th t1 row of the grid, "
                                    f"extrapolated x{w['stride']} (one full N=20k permutation costs about a minute of one core)")},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def load_ncu_captures(cfg_name):
    """Per-kernel static counters of this build (tools/ncu_capture.sh -> profiles/kernel_counters_<cfg>.json).  A capture
    taken from other kernel sources is refused: instruction counts only describe the build they were measured on."""
    path = os.path.join(ROOT, "profiles", f"kernel_counters_{cfg_name}.json")
    try:
        d = json.load(open(path))
    except Exception:
        return None, f"no capture at profiles/kernel_counters_{cfg_name}.json"
    if d.get("kernel_source_sha") != kernel_source_sha():
        return None, (f"profiles/kernel_counters_{cfg_name}.json was captured from kernel sources {d.get('kernel_source_sha')}, "
                      f"this build is {kernel_source_sha()}: refused")
    return d, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--perms", type=int, default=0, help="override the permutations per step of the workload")
    ap.add_argument("--pairs", type=int, default=0, help="c4: override the number of list pairs")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-row-stride", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--sigma-ctas", type=int, default=0, help="pairing kernel CTAs per SM: 0 = library default, 1 or 2 = force (A/B)")
    ap.add_argument("--cpu-worker", default="", choices=["", "baseline", "reference"], help=argparse.SUPPRESS)
    args = ap.parse_args()

    if args.cpu_worker:
        cpu_worker(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
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
    from dual_threshold_optimization_b200 import _capi as capi
    from dual_threshold_optimization_b200.run import run_pairs_structs, run_single_node_records
    from dual_threshold_optimization_b200.stat_operations import empirical_pvalue_struct
    import ctypes as C

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    cfg = CONFIGS[args.config]
    eng = dto.Engine(local_rank)  # raises without a GPU: there is no CPU path to fall back to
    if args.batch:
        eng.set_option("batch", args.batch)
    if args.sigma_ctas:
        eng.set_option("sigma_ctas", args.sigma_ctas)
    clocks = ClockSampler(local_rank)

    # the product's own NCCL all-gather (dto_b200_allgather_minima): rank 0 makes the unique id, torch's store carries it
    comm = C.c_void_p()
    if world > 1:
        uid = (C.c_char * 128)()
        if rank == 0:
            capi.check(capi.lib().dto_b200_nccl_unique_id(uid))
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0)
        uid = (C.c_char * 128).from_buffer_copy(box[0])
        capi.check(capi.lib().dto_b200_nccl_comm_create(C.byref(comm), world, uid, rank, local_rank))

    line = None
    if args.config == "c4":
        line = bench_pairs(args, cfg, dto, eng, clocks, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks, run_pairs_structs)
    else:
        line = bench_single(args, cfg, dto, capi, eng, clocks, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks,
                            run_single_node_records, empirical_pvalue_struct, comm, torch, dist)
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        if comm:
            capi.lib().dto_b200_nccl_comm_destroy(comm)
        dist.destroy_process_group()


def shard(total, world, rank):
    """contiguous ceil-chunks, like multi_node.rs:117-120"""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def bench_single(args, cfg, dto, capi, eng, clocks, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks,
                 run_single_node_records, empirical_pvalue_struct, comm, torch, dist):
    import ctypes as C

    ids1, r1, ids2, r2, bg = workload_lists(args.config)
    l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
    population = dto.compute_population_size(l1, l2, dto.FeatureList(bg) if bg is not None else None)
    eng.load_lists(l1, l2, population)
    unperm = eng.run_unpermuted()
    T1, T2 = eng.shape[0], eng.shape[1]
    if args.scaling == "strong":
        total = args.perms or cfg.get("perms_strong", cfg["perms"])
        lo, hi = shard(total, world, rank)
        P = hi - lo          # this rank's permutations per step
        P_max = shard(total, world, 0)[1]
        per_step_all = total
    else:
        P = P_max = args.perms or cfg["perms"]
        lo = rank * P
        per_step_all = P * world
    dev = f"cuda:{local_rank}"
    d_minp = torch.zeros(max(P_max, 1), dtype=torch.float64, device=dev)
    gathered = torch.zeros(max(P_max, 1) * world, dtype=torch.float64, device=dev) if world > 1 else d_minp

    def device_step(step_idx):
        first = 1 + step_idx * per_step_all + lo  # disjoint id ranges across ranks and steps
        if P:
            eng.run_permuted_philox_device(PHILOX_SEED, first, P, d_minp.data_ptr())
        ms = eng.stats()["last_run_ms"] if P else 0.0
        if world > 1:  # the small all-gather of per-permutation minima over NVLink, issued by the library itself
            capi.check(capi.lib().dto_b200_allgather_minima(eng.ctx, comm, C.c_void_p(d_minp.data_ptr()), C.c_void_p(gathered.data_ptr()), P_max))
        return ms

    if rank == 0:
        clocks.start()  # started before the warm-up (nvidia-smi needs a few 100 ms to come up); marked at the timed start
    for w in range(args.warmup):
        device_step(10_000 + w)
    barrier()
    eng.reset_stats()
    barrier()
    clocks.mark()
    t_wall0 = time.perf_counter()
    dev_ms = scan_ms = sigma_ms = 0.0
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
    emp = float((gathered <= float(unperm["pvalue"])).double().mean().item())

    dev_ms_max, wall_ms_max = max_over_ranks([dev_ms, wall_ms])
    value = per_step_all * args.steps / (wall_ms_max * 1e-3)  # whole job, barrier-to-barrier, max over ranks
    value_device_only = per_step_all * args.steps / (dev_ms_max * 1e-3)
    tie_tasks, full_tasks = sum_over_ranks([st["tasks_tie_resolved"], st["tasks_full"]])

    # ---- end to end through the drop-in boundary: dto_b200_run_tasks on list handles (host strings -> slot map -> H2D,
    # unpermuted task on rank 0, this rank's id range of permuted tasks, records -> host), records of all ranks gathered
    # (NCCL, like the MPI gather of multi_node.rs:148-160), dto_b200_empirical_pvalue on rank 0 ----
    n_un = 1 if rank == 0 else 0
    ids_arr = np.zeros(n_un + P, dtype=np.uint64)
    perm_arr = np.ones(n_un + P, dtype=np.uint8)
    if n_un:
        perm_arr[0] = 0
    rec_bytes = capi.RECORD_DTYPE.itemsize
    # host result buffer of one step: [unpermuted record, permuted records of every rank]; pinned, so the device copies
    # land in it directly and dto_b200_empirical_pvalue reads it in place
    host_all = torch.zeros((1 + world * P_max) * rec_bytes, dtype=torch.uint8).pin_memory()
    all_recs = host_all.numpy().view(capi.RECORD_DTYPE)
    my_recs = all_recs[: n_un + P] if world == 1 else np.zeros(n_un + P, dtype=capi.RECORD_DTYPE)
    stage_in = torch.zeros(P_max * rec_bytes, dtype=torch.uint8).pin_memory() if world > 1 else None
    gather_in = torch.zeros(P_max * rec_bytes, dtype=torch.uint8, device=dev) if world > 1 else None
    gather_out = torch.zeros(world * P_max * rec_bytes, dtype=torch.uint8, device=dev) if world > 1 else None
    uneven = args.scaling == "strong" and per_step_all % world != 0
    e2e_copy_bytes = [0, 0]

    def e2e_step(step_idx):
        first = 1 + (50_000 + step_idx) * per_step_all + lo
        ids_arr[n_un:] = np.arange(first, first + P, dtype=np.uint64)
        recs = run_single_node_records((ids_arr, perm_arr), l1, l2, population, 1, [local_rank], PHILOX_SEED, out=my_recs)
        if world == 1:
            return empirical_pvalue_struct(recs).empirical_pvalue
        # the records of all ranks meet on rank 0 (NCCL all-gather of the raw records: what multi_node.rs:148-160 does over MPI)
        raw = recs[n_un:].view(np.uint8).ravel()
        stage_in[: raw.size] = torch.from_numpy(raw)
        gather_in.copy_(stage_in, non_blocking=True)
        dist.all_gather_into_tensor(gather_out, gather_in)
        e2e_copy_bytes[0] += gather_in.numel()
        if rank != 0:
            torch.cuda.synchronize()
            return None
        host_all[rec_bytes:].copy_(gather_out, non_blocking=True)
        torch.cuda.synchronize()
        e2e_copy_bytes[1] += gather_out.numel()
        all_recs[0] = recs[0]
        if not uneven:
            return empirical_pvalue_struct(all_recs).empirical_pvalue
        parts = [all_recs[:1]]
        for r in range(world):
            rlo, rhi = shard(per_step_all, world, r)
            parts.append(all_recs[1 + r * P_max: 1 + r * P_max + (rhi - rlo)])
        return empirical_pvalue_struct(np.concatenate(parts)).empirical_pvalue

    e2e_step(-1)
    barrier()
    tot0 = capi.process_totals()
    e2e_copy_bytes[0] = e2e_copy_bytes[1] = 0
    t0 = time.perf_counter()
    e2e_emp = None
    for s in range(args.e2e_steps):
        e2e_emp = e2e_step(s)
    barrier()
    e2e_s = time.perf_counter() - t0
    (e2e_s_max,) = max_over_ranks([e2e_s])
    e2e_value = per_step_all * args.e2e_steps / e2e_s_max
    # bytes moved per step on this rank, counted by the library at every copy it issues (lists, slot map, tables' inputs up;
    # records, summaries down) plus the gather staging of the multi-rank run
    tot1 = capi.process_totals()
    n1 = len(ids1)
    h2d_step = (tot1["h2d_bytes"] - tot0["h2d_bytes"] + e2e_copy_bytes[0]) // max(args.e2e_steps, 1)
    d2h_step = (tot1["d2h_bytes"] - tot0["d2h_bytes"] + e2e_copy_bytes[1]) // max(args.e2e_steps, 1)
    e2e_launches = tot1["launches"] - tot0["launches"]

    # ---- the same job through ONE process driving all GPUs: dto_b200_run_tasks(devices = [0 .. N-1]) -- what `-m` of the CLI
    # maps to, the single-box replacement of src/run/multi_node.rs:114-161.  Rank 0 alone runs it (the other ranks wait at
    # the barrier with idle GPUs); reported next to the process-per-GPU figure, not instead of it. ----
    in_process = None
    if world > 1:
        store = dist.distributed_c10d._get_default_store()
        torch.cuda.synchronize()
        if rank == 0:
            ids_all = np.zeros(1 + per_step_all, dtype=np.uint64)
            perm_all = np.ones(1 + per_step_all, dtype=np.uint8)
            perm_all[0] = 0
            out_all = np.zeros(1 + per_step_all, dtype=capi.RECORD_DTYPE)
            devs = list(range(world))

            def in_process_step(step_idx):
                first = 1 + (70_000 + step_idx) * per_step_all
                ids_all[1:] = np.arange(first, first + per_step_all, dtype=np.uint64)
                recs = run_single_node_records((ids_all, perm_all), l1, l2, population, 1, devs, PHILOX_SEED, out=out_all)
                return empirical_pvalue_struct(recs).empirical_pvalue

            in_process_step(-1)
            t0 = time.perf_counter()
            for s in range(args.e2e_steps):
                in_process_step(s)
            dt = time.perf_counter() - t0
            in_process = {"value": per_step_all * args.e2e_steps / dt, "unit": "permutations/s", "steps": args.e2e_steps, "devices": devs,
                          "path": "one process, dto_b200_run_tasks(devices = all GPUs) + dto_b200_empirical_pvalue; host threads, no NCCL"}
            store.set("dto_in_process_done", "1")
        else:
            # wait on the CPU (the rendezvous store), not in an NCCL barrier: a spinning NCCL kernel of this process would
            # time-slice this GPU against rank 0's kernels
            store.wait(["dto_in_process_done"])
        barrier()

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        perms_timed = P * args.steps
        clock_hz = ((clk or {}).get("sm_mhz") or 1965.0) * 1e6
        n1_eff = int(np.searchsorted(l1.ranks(), l1.thresholds()[-1], side="right"))
        pb_row_bytes = ((n1_eff + 1 + 255) // 256) * 256 * 2
        ncu, why = load_ncu_captures(args.config)
        issue_peak = 148 * 4 * clock_hz
        n_common = n1  # identical gene sets in every single-pair workload here
        alg_int_inst = (n_common + 2.0 * T1 * T2) / 32.0  # N shared-memory atomics + T1*T2 integer adds (+ as many compares), per warp instruction
        kernels = []
        for name, ms_tot, key in (("dto::scan_kernel", scan_ms, "scan"), ("dto::sigma_sort_kernel", sigma_ms, "sigma")):
            k = {"kernel": name, "ms_per_step": ms_tot / args.steps, "share_of_step": ms_tot / dev_ms if dev_ms else None}
            if ncu and key in ncu:
                wi = ncu[key]["warp_instructions_per_permutation"]
                k.update({"warp_instructions_per_permutation": wi,
                          "issue_utilisation": wi * perms_timed / (ms_tot * 1e-3) / issue_peak if ms_tot else None,
                          "ncu_issue_active_pct": ncu[key]["issue_active_pct"],
                          "dram_bytes_per_permutation": ncu[key]["dram_bytes_per_permutation"]})
            kernels.append(k)
        hbm_alg_bytes = 2 * pb_row_bytes + 40  # pairing writes the row, the scan reads it, one record out
        hbm_view = {"bound": "hbm", "achieved": hbm_alg_bytes * perms_timed / (dev_ms * 1e-3) / 1e9 if dev_ms else None,
                    "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "algorithmic_bytes_per_permutation": hbm_alg_bytes,
                    "note": "partner-slot row written by the pairing kernel and read by the scan + one 40 B record; peak = MEASURED_PEAKS.json hbm_gbs (of measured)"}
        if hbm_view["achieved"] and hbm_view["peak"]:
            hbm_view["frac"] = hbm_view["achieved"] / hbm_view["peak"]
        if ncu:
            wi_step = sum(ncu[k]["warp_instructions_per_permutation"] for k in ("scan", "sigma") if k in ncu)
            inst_rate = wi_step * perms_timed / (dev_ms * 1e-3)
            roof = {"bound": "issue", "achieved": inst_rate / 1e9, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s", "frac": inst_rate / issue_peak,
                    "traffic": sum(ncu[k]["dram_bytes_per_permutation"] for k in ("scan", "sigma") if k in ncu) * (perms_timed / max(scan_launches, 1)),
                    "warp_instructions_per_permutation": wi_step,
                    "algorithmic_integer_fraction": alg_int_inst / wi_step,
                    "algorithmic_warp_instructions_per_permutation": alg_int_inst,
                    "kernel_source_sha": ncu["kernel_source_sha"], "capture": ncu.get("capture"),
                    "note": ("whole step (pairing + scan): warp instructions per permutation from the ncu capture of THIS build (source hash checked) x live "
                             "permutation rate (CUDA events on the library's stream) vs 148 SMs x 4 schedulers x SM clock.  Neither 'hbm' nor 'tensor' binds this "
                             "integer/byte path: FP64 pipe ~2 %, DRAM ~5 % busy (ncu).  algorithmic_integer_fraction = SURVEY 8(d)'s integer work (n_common "
                             "atomics + 2 T1 T2 adds/compares, as warp instructions) / instructions actually issued.")}
        else:
            roof = dict(hbm_view)
            roof["note"] = f"no valid ncu capture for this build ({why}); algorithmic HBM view only. " + hbm_view["note"]
            roof["traffic"] = None
        roof["kernels"] = kernels
        roof["step_kernel_ms_sum"] = (scan_ms + sigma_ms) / args.steps
        roof["ms_per_step_device"] = dev_ms / args.steps
        roof["launches_per_step"] = {"scan": scan_launches / args.steps, "pairing_sort": scan_launches / args.steps}
        roof["hbm_view"] = hbm_view
        screened = st["level2_cells"] / max(st["tasks_fast"], 1)
        roof["post_pruning"] = {"cells_per_permutation": T1 * T2, "cells_past_screen_per_permutation": screened,
                                "pruning_rate": 1.0 - screened / (T1 * T2),
                                "exact_tail_evaluations_per_permutation": st["candidates"] / max(st["tasks_fast"], 1)}
        cpu = None
        if not args.no_cpu_baseline:
            w = run_cpu_worker("baseline", args)
            fp64_peak = eng.probe_fp64_tflops()
            f_perm = w["algorithmic_flops_per_permutation"]
            ach = f_perm * perms_timed / (scan_ms * 1e-3) / 1e12 if scan_ms else None
            roof["fp64_view"] = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if ach else None,
                                 "algorithmic_flops_per_permutation": f_perm,
                                 "note": ("SURVEY 8(d) convention: 26 + 5R FP64 flops per evaluated cell (R = oracle-counted converged tail on the same input). The scan "
                                          "certifies-and-skips almost every cell, so this algorithmic rate exceeds the pipe peak by design; peak = DFMA chain measured live "
                                          "(MEASURED_PEAKS.json holds no FP64 figure).")}
            cpu = {"value": w["faithful_value"], "unit": "permutations/s", "cores": w["cores"], "kind": "port",
                   "sample": (f"reference-faithful oracle mode (string ids, per-cell hash set, uncached ln_gamma, full tails), {w['cores']} permutations on "
                              f"{w['cores']} threads, every {w['stride']}th t1 row of the grid, extrapolated x{w['stride']}; took {w['faithful_seconds']:.1f} s"),
                   "optimized_port_value": w["optimized_value"],
                   "optimized_port_sample": (f"integer-id oracle mode (histogram + prefix sum, cached ln-factorial), {w['optimized_perms']} full permutations on "
                                             f"{w['cores']} threads in {w['optimized_seconds']:.1f} s")}
        line = {
            "metric": METRIC, "value": value, "unit": "permutations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args, args.config), "features": n1, "threshold_pairs": T1 * T2,
                       "permutations_per_step": per_step_all, "permutations_per_gpu_per_step": P_max,
                       "l2_policy": f"inputs larger than L2: each step streams {P_max * pb_row_bytes / 1e9:.2f} GB of partner-slot rows per GPU (126 MB L2)"
                                    if P_max * pb_row_bytes > 126e6 else "inputs smaller than L2: every step draws fresh permutation ids (new rows), tables stay resident by design",
                       "pvalue_evals_per_s": value * T1 * T2},
            "device_only_value": value_device_only,
            "e2e": {"value": e2e_value, "unit": "permutations/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                    "steps": args.e2e_steps, "empirical_pvalue": e2e_emp, "gpu_launches": int(e2e_launches), "in_process_all_gpus": in_process,
                    "path": "dto_b200_run_tasks (list handles -> records on the host) + dto_b200_empirical_pvalue" + (", records gathered over NCCL" if world > 1 else "")},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_ms": {"scan": scan_ms / args.steps, "pairing_sort": sigma_ms / args.steps, "device_total": dev_ms / args.steps},
            "near_tie_rate": tie_tasks / max(per_step_all * args.steps, 1), "tasks_full": int(full_tasks),
            "stats": {k: st[k] for k in ("tasks_fast", "tasks_full", "tasks_tie_resolved", "tie_cells_host", "candidates", "level2_cells", "refined_cells")},
            "unpermuted": {"rank1": int(unperm["rank1"]), "rank2": int(unperm["rank2"]), "pvalue": float(unperm["pvalue"]), "empirical_pvalue": emp},
        }
    return line


def pairs_roofline(tasks_per_s, sm_mhz):
    """c4 runs the N = 6000 instantiations of the two kernels (the c2 capture) inside a batched driver: whole-job issue
    utilisation = warp instructions per task x tasks/s (wall clock, host work included) / issue peak."""
    ncu, why = load_ncu_captures("c2")
    peak = 148 * 4 * sm_mhz * 1e6
    if not ncu:
        return {"bound": "issue", "achieved": None, "peak": peak / 1e9, "unit": "Gwarp-inst/s", "frac": None, "traffic": None,
                "note": f"no valid ncu capture of the N = 6000 kernels for this build ({why})"}
    wi = ncu["scan"]["warp_instructions_per_permutation"] + ncu["sigma"]["warp_instructions_per_permutation"]
    rate = wi * tasks_per_s
    return {"bound": "issue", "achieved": rate / 1e9, "peak": peak / 1e9, "unit": "Gwarp-inst/s", "frac": rate / peak,
            "traffic": (ncu["scan"]["dram_bytes_per_permutation"] + ncu["sigma"]["dram_bytes_per_permutation"]) * tasks_per_s,
            "warp_instructions_per_permutation": wi, "kernel_source_sha": ncu["kernel_source_sha"],
            "note": ("whole job through dto_b200_run_pairs on the wall clock (host canonicalisation, uploads, epilogue included): warp instructions per task "
                     "of the N = 6000 kernels (ncu capture of this build, profiles/kernel_counters_c2.json) x tasks/s vs 148 SMs x 4 schedulers x SM clock; "
                     "`traffic` is DRAM bytes per second here, not per launch (launch sizes vary with the grouping)")}


def bench_pairs(args, cfg, dto, eng, clocks, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks, run_pairs_structs):
    """c4: the batch of list pairs, sharded by pair over ranks; every step = the whole batch through dto_b200_run_pairs
    (per pair: slot map from string ids, lists up, unpermuted task + 1 000 Philox permutations, records down, epilogue)."""
    n_pairs = args.pairs or cfg["pairs"]
    P = args.perms or cfg["perms_per_pair"]
    lo, hi = shard(n_pairs, world, rank)
    t0 = time.perf_counter()
    pairs = []
    for q in range(lo, hi):
        ids1, r1, ids2, r2 = S.synthetic_pair(cfg["features"], 1 + q, (0.25, None, 0.35, None)[q % 4])
        l1, l2 = dto.RankedFeatureList.from_(ids1, r1), dto.RankedFeatureList.from_(ids2, r2)
        pairs.append((l1, l2, cfg["features"]))
    build_s = time.perf_counter() - t0
    seed_rank = (PHILOX_SEED + lo * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF  # pair q keeps the seed of the unsharded run

    def step(s):
        return run_pairs_structs(pairs, P, devices=[local_rank], seed=(seed_rank + s * 7919) & 0xFFFFFFFFFFFFFFFF)

    from dual_threshold_optimization_b200 import _capi as capi

    if rank == 0:
        clocks.start()
    for w in range(args.warmup):
        step(100 + w)
    barrier()
    clocks.mark()
    tot0 = capi.process_totals()
    t0 = time.perf_counter()
    out = None
    for s in range(args.steps):
        out = step(s)
    barrier()
    wall = time.perf_counter() - t0
    tot1 = capi.process_totals()
    clk = clocks.stop() if rank == 0 else None
    (wall_max,) = max_over_ranks([wall])
    pairs_per_s = n_pairs * args.steps / wall_max
    value = pairs_per_s * P
    T = int(pairs[0][0].thresholds().size) if pairs else 0
    line = None
    if rank == 0:
        n = cfg["features"]
        cpu = None
        if not args.no_cpu_baseline:
            w = run_cpu_worker("baseline", args)
            cpu = {"value": w["faithful_value"], "unit": "permutations/s", "cores": w["cores"], "kind": "port",
                   "sample": (f"reference-faithful oracle mode on one N=6000 pair, {w['cores']} permutations on {w['cores']} threads, every "
                              f"{w['stride']}th t1 row, extrapolated x{w['stride']}; took {w['faithful_seconds']:.1f} s"),
                   "optimized_port_value": w["optimized_value"]}
        sig = sum(1 for x in out if x.empirical_pvalue <= 0.01)
        line = {
            "metric": METRIC, "value": value, "unit": "permutations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_string(args, args.config), "features": n, "threshold_pairs": T * T, "pairs": n_pairs,
                       "permutations_per_pair": P, "permutations_per_step": n_pairs * P, "pairs_per_s": pairs_per_s, "pvalue_evals_per_s": (value + pairs_per_s) * T * T,
                       "l2_policy": "every pair brings new lists and every step new permutation ids; ~1.5 GB of partner-slot rows per launch (126 MB L2)",
                       "list_build_s_rank0": build_s},
            "e2e": {"value": value, "unit": "permutations/s", "h2d_bytes_per_step": int((tot1["h2d_bytes"] - tot0["h2d_bytes"]) // args.steps),
                    "d2h_bytes_per_step": int((tot1["d2h_bytes"] - tot0["d2h_bytes"]) // args.steps), "steps": args.steps,
                    "path": "dto_b200_run_pairs on list handles (one context per GPU): value and e2e are the same call here -- every pair's lists are new host data by definition"},
            "gpu_launches": int(tot1["launches"] - tot0["launches"]), "clocks": clk,
            "roofline": pairs_roofline(value + pairs_per_s, (clk or {}).get("sm_mhz") or 1965.0),
            "cpu_baseline": cpu,
            "pairs_significant_at_0.01": sig,
        }
    return line


if __name__ == "__main__":
    main()
