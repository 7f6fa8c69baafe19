#!/usr/bin/env python3
"""Condenses an `ncu --set full` report into the small JSON that is committed under profiles/ and that
bench.py reads for `roofline.traffic`.  Usage: python profiles/ncu_summary.py REPORT.ncu-rep OUT.json "launch description" ALG_BYTES"""
import csv
import json
import subprocess
import sys

KEYS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size",
    "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out, launch, alg = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        m = {k: {"value": r[head.index(k)], "unit": units[head.index(k)]} for k in KEYS if k in head}
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tr = sum(float(m[k]["value"]) * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        res.append({"kernel": r[head.index("Kernel Name")], "launch": launch, "metrics": m,
                    "traffic_bytes_per_launch": int(tr), "algorithmic_bytes_per_launch": alg})
    json.dump(res[0] if len(res) == 1 else res, open(out, "w"), indent=1)
    for x in res:
        print(x["kernel"][:60], x["traffic_bytes_per_launch"], x["metrics"]["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()
