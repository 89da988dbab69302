#!/usr/bin/env python3
"""Summarise `ncu --page raw --csv` exports (one kernel launch per file) into the handful of numbers DESIGN.md and
profiles/ quote.  usage: scripts/ncu_summary.py gpurun_out/prof/raw_*.csv"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct",
    "smsp__issue_active.avg.pct": "issue_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe_pct",
    "launch__registers_per_thread": "regs",
    "launch__occupancy_limit_shared_mem": "occ_lim_smem",
    "launch__occupancy_limit_registers": "occ_lim_regs",
    "launch__grid_size": "grid",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_per_cycle",
    "smsp__warps_active.avg.per_cycle_active": "warps_per_sched",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__inst_executed.sum": "inst_executed",
}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    if hdr is None:
        return None
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    d = dict(zip(names, vals))
    u = dict(zip(names, units))
    out = {"kernel": d.get("Kernel Name", "?")[:90]}
    for k, nm in KEYS.items():
        if k in d:
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            unit = u.get(k, "")
            if nm == "time_us":
                v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
            if nm in ("dram_read", "dram_write", "l2_bytes"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                v *= mult
            out[nm] = v
    stalls = {k.split("smsp__pcsamp_warps_issue_stalled_")[1]: float(v.replace(",", "")) for k, v in d.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "n/a")}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
    out["stalls"] = top
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        s = load(p)
        print(p)
        print(json.dumps(s, indent=None))
