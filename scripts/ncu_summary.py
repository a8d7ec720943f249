#!/usr/bin/env python
"""Summarise an `ncu --set full` report for profiles/: one column per captured kernel launch, the metrics the
DESIGN / bench rooflines quote (duration, DRAM bytes, pipe utilisation, shared-memory wavefronts, top stalls).
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv [profiles/x.json]"""
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    names = [dict(zip(hdr, r))["Kernel Name"].split("(")[0] for r in data]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in METRICS + stalls:
            if m in hdr:
                k = hdr.index(m)
                w.writerow([m, units[k]] + [r[k] for r in data])
    if len(sys.argv) > 3:
        js = {}
        for r in data:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d["Kernel Name"].split("(")[0]
            dur = float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]]
            js[name] = {
                "duration_ms": dur,
                "dram_bytes": to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) +
                              to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]),
                "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                "fp64_pipe_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
                "shared_wavefronts": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", "")),
                "shared_wavefronts_pct_of_peak": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]),
                "l1tex_pct": float(d["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]),
                "inst_executed": float(d["smsp__inst_executed.sum"].replace(",", "")),
            }
        json.dump(js, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
