#!/usr/bin/env python
"""One-paragraph summary per `ncu --set full` raw CSV (profiles/r02_ncu_full_*_raw.csv): duration, DRAM bytes, occupancy,
pipe utilisation and the dominant warp stall reasons, averaged over the captured launches."""
import csv
import glob
import os
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__occupancy_limit_registers", "occ_lim_regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cycles_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("smsp__inst_executed.sum", "inst"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct")]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def summarize(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    if hdr is None:
        return f"{os.path.basename(path)}: no kernel rows"
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    col = {n: i for i, n in enumerate(names)}
    kname = data[0][col["Kernel Name"]] if data else "?"
    out = [f"{os.path.basename(path)}: {kname[:60]}  ({len(data)} launches)"]
    vals = {}
    for key, lab in KEYS:
        if key in col:
            xs = [num(r[col[key]]) for r in data]
            xs = [x for x in xs if x is not None]
            if xs:
                vals[lab] = (sum(xs) / len(xs), units[col[key]])
    def show(lab, scale=1.0, fmt="{:.4g}"):
        if lab in vals:
            return f"{lab}=" + fmt.format(vals[lab][0] * scale) + (f" {vals[lab][1]}" if scale == 1.0 else "")
        return None
    parts = [show(l) for l in ("duration", "dram_read", "dram_write", "regs", "grid", "block", "warps_active_pct", "dram_pct",
                                "fp64_pipe_pct", "fp64_cycles_pct", "lsu_pct", "sm_pct", "l1_hit_pct", "l2_hit_pct",
                                "smem_bank_conflicts", "inst")]
    out.append("   " + "; ".join(p for p in parts if p))
    # warp stall reasons (pcsamp issue-stall ratios)
    stalls = []
    for n, i in col.items():
        if n.startswith("smsp__average_warp_latency_issue_stalled_") or n.startswith("smsp__average_warps_issue_stalled_"):
            xs = [num(r[i]) for r in data]
            xs = [x for x in xs if x is not None]
            if xs:
                stalls.append((sum(xs) / len(xs), n.split("stalled_")[1].replace("_per_warp_active.pct", "").replace(".ratio", "")))
    stalls.sort(reverse=True)
    if stalls:
        out.append("   top stalls: " + ", ".join(f"{n} {v:.3g}" for v, n in stalls[:5]))
    return "\n".join(out)


if __name__ == "__main__":
    files = sys.argv[1:] or sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_full_*_raw.csv")))
    for f in files:
        print(summarize(f))
