"""Turns an .ncu-rep (read with the local ncu, no GPU needed) into the text summary committed under profiles/."""
import csv
import io
import subprocess
import sys
import collections

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALLS = "smsp__average_warps_issue_stalled_"


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main(rep, tasks):
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# ncu summary of {rep.split('/')[-1]}  (ncu --set full --clock-control none; one launch, {tasks} permutations)")
    print(f"kernel: {m.get('Kernel Name', ('?',))[0]}")
    for k in KEYS:
        if k in m:
            print(f"{k:75s} {m[k][0]:>18s} {m[k][1]}")
    print("\n# warp stall reasons (warps per issue-active cycle)")
    for h in hdr:
        if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
            print(f"{h[len(STALLS):-len('_per_issue_active.ratio')]:30s} {m[h][0]}")
    inst = float(m["smsp__inst_executed.sum"][0])
    dur_us = float(m["gpu__time_duration.sum"][0]) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(m["gpu__time_duration.sum"][1], 1.0)
    print(f"\nwarp instructions per permutation: {inst / tasks:.0f}; kernel time per permutation: {dur_us / tasks:.3f} us "
          f"(under ncu: cold caches, serialised)")
    print(f"DRAM bytes per permutation: read {float(m['dram__bytes_read.sum'][0]) * {'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}.get(m['dram__bytes_read.sum'][1],1) / tasks:.0f}, "
          f"write {float(m['dram__bytes_write.sum'][0]) * {'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}.get(m['dram__bytes_write.sum'][1],1) / tasks:.0f}")
    # source-level hot spots
    rows = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur = None
    agg, src = collections.OrderedDict(), {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name"):
            continue
        if r[2] == "-" and r[0].isdigit():
            key = (cur, int(r[0]))
            a = agg.setdefault(key, [0, 0])
            a[0] += int(r[6] or 0)
            a[1] += int(r[7] or 0)
            src[key] = r[1]
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print("\n# top source lines by warp-stall samples (share of samples, share of executed warp instructions)")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{key[0][:16]:16s}:{key[1]:4d}  samp {100 * a[0] / ts:5.1f}%  inst {100 * a[1] / ti:5.1f}%  {src[key].strip()[:100]}")


def to_json(rep, tasks, out_path):
    """Key per-launch figures of one captured launch, for bench.py's roofline object (static evidence, not live)."""
    import json

    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def f(k):
        return float(m[k][0])

    d = {
        "source": rep.split("/")[-1], "permutations_in_launch": tasks,
        "dram_bytes_per_permutation": (f("dram__bytes_read.sum") * scale.get(m["dram__bytes_read.sum"][1], 1) +
                                       f("dram__bytes_write.sum") * scale.get(m["dram__bytes_write.sum"][1], 1)) / tasks,
        "warp_instructions_per_permutation": f("smsp__inst_executed.sum") / tasks,
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
        "alu_pipe_pct": f("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
        "shared_wavefronts_pct_of_peak": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "shared_atomic_instructions_per_permutation": f("smsp__inst_executed_op_shared_atom.sum") / tasks,
        "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "top_stalls": {h[len(STALLS):-len("_per_issue_active.ratio")]: float(m[h][0]) for h in hdr
                       if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and float(m[h][0]) > 0.2 and "selected" not in h},
    }
    json.dump(d, open(out_path, "w"), indent=1)


def counters(cfg, tag, tasks):
    """profiles/kernel_counters_<cfg>.json (+ text summaries) from gpurun_out/<tag>_{scan,sigma}_<cfg>.ncu-rep."""
    import contextlib
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    go = os.path.join(root, "gpurun_out")
    sha = open(os.path.join(go, f"{tag}_{cfg}_sha.txt")).read().strip()
    out = {"kernel_source_sha": sha,
           "capture": {"tool": "ncu --set full --clock-control none --import-source on, one launch per kernel (4th launch: after 3 warm-up steps)",
                       "command": f"bash tools/ncu_capture.sh {cfg} {tag} {tasks}", "permutations_in_launch": tasks, "tag": tag}}
    for k in ("scan", "sigma"):
        rep = os.path.join(go, f"{tag}_{k}_{cfg}.ncu-rep")
        tmp = os.path.join(go, f"{tag}_{k}_{cfg}.json")
        to_json(rep, tasks, tmp)
        out[k] = json.load(open(tmp))
        with open(os.path.join(root, "profiles", f"{tag}_{k}_kernel_ncu_{cfg}.txt"), "w") as f, contextlib.redirect_stdout(f):
            main(rep, tasks)
    json.dump(out, open(os.path.join(root, "profiles", f"kernel_counters_{cfg}.json"), "w"), indent=1)
    print(json.dumps({k: {x: out[k][x] for x in ("warp_instructions_per_permutation", "issue_active_pct", "dram_bytes_per_permutation")} for k in ("scan", "sigma")}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "counters":
        counters(sys.argv[2], sys.argv[3], int(sys.argv[4]))
    elif len(sys.argv) > 3:
        to_json(sys.argv[1], int(sys.argv[2]), sys.argv[3])
    else:
        main(sys.argv[1], int(sys.argv[2]))
