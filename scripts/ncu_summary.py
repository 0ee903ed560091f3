#!/usr/bin/env python
"""Summarise `ncu --set full` reports for profiles/ and refresh profiles/ncu_counters.json (what bench.py reads for
`roofline.traffic` and `fp64_pipe_pct`).

    python scripts/ncu_summary.py <tag> <report.ncu-rep> <key> [<report2> <key2> ...]

<key> names the kernel AND the size it was captured at, e.g. "k_rk4_rollout<1,0,0>@1048576x1000"; bench.py looks the key
up with the sizes it runs.  Writes profiles/<tag>_<kernel>_ncu_summary.txt and updates profiles/ncu_counters.json.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio$")
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def raw_rows(rep):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    r = list(csv.reader(io.StringIO(txt)))
    names, units = r[0], r[1]
    return [dict(zip(names, zip(units, row))) for row in r[2:]]


def main():
    tag = sys.argv[1]
    pairs = list(zip(sys.argv[2::2], sys.argv[3::2]))
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    try:
        db = json.load(open(path))
    except Exception:
        db = {"what": "counters from `ncu --set full --clock-control none` captures of the product kernels at the benchmark's own sizes "
                      "(scripts/ncu_summary.py); bench.py reports dram_bytes as roofline.traffic and fp64_pipe_pct next to roofline.frac",
              "kernels": {}}
    for rep, key in pairs:
        rows = raw_rows(rep)
        kname = key.split("@")[0].split("<")[0]
        row = [r for r in rows if kname in r["Kernel Name"][1]][-1]
        short = re.sub(r"[^A-Za-z0-9]+", "_", key).strip("_")       # kernel + template arguments + size: one file per key
        out = os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (tag, short))
        lines = ["ncu --set full --clock-control none --import-source on; key %s; report %s" % (key, os.path.basename(rep)), "-----",
                 "Kernel Name = %s" % row["Kernel Name"][1], "Block Size = %s" % row["Block Size"][1], "Grid Size = %s" % row["Grid Size"][1]]
        for k in KEEP:
            if k in row:
                lines.append("%s = %s %s" % (k, row[k][1], row[k][0]))
        for k in sorted(row):
            if STALL.search(k):
                lines.append("%s = %s" % (k, row[k][1]))
        open(out, "w").write("\n".join(lines) + "\n")

        def val(k):
            u, v = row[k]
            return float(v.replace(",", "")) * UNIT.get(u, 1.0)
        db["kernels"][key] = {
            "kernel": row["Kernel Name"][1], "gpu_time_ms": float(row["gpu__time_duration.sum"][1]) * {"ms": 1, "us": 1e-3, "s": 1e3}.get(row["gpu__time_duration.sum"][0], 1),
            "fp64_pipe_pct": float(row["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"][1]),
            "dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
            "registers": int(float(row["launch__registers_per_thread"][1])),
            "local_loads": int(float(row["sass__inst_executed_local_loads"][1])),
            "local_stores": int(float(row["sass__inst_executed_local_stores"][1])),
            "smem_bank_conflicts": int(float(row["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][1])),
            "source": "profiles/" + os.path.basename(out),
        }
        print(open(out).read())
    json.dump(db, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
