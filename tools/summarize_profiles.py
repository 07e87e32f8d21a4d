#!/usr/bin/env python
"""Turn the ncu reports of tools/collect_profiles.sh (gpurun_out/*.ncu-rep, scratch) into the tracked text summaries
and profiles/ncu_traffic.json (DRAM bytes per launch, the `roofline.traffic` of bench.py).  Needs ncu, no GPU."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    gp = os.path.join(ROOT, "gpurun_out")
    traffic = {}
    for workload, reps in (("C4", ["r2_stats_full_c4", "r2_sweep_c4"]), ("C3", ["r2_stats_full_c3", "r2_sweep_c3"])):
        lines = ["ncu --set full --clock-control none, bench.py --workload %s --steps 3 --warmup 3 --no-e2e --no-cpu "
                 "--no-parity (tools/collect_profiles.sh); the first statistic launch of a run is a forced FULL pass, "
                 "the later ones are the incremental correction" % workload, ""]
        traffic[workload] = {}
        for rep in reps:
            path = os.path.join(gp, rep + ".ncu-rep")
            if not os.path.exists(path):
                continue
            launches, units = rows_of(path)
            for d in launches:
                name = d["Kernel Name"]
                short = "k_cells_sell" if "k_cells_sell" in name else ("k_table" if "k_table" in name else "k_stats")
                rd = float(d["dram__bytes_read.sum"]) * SCALE[units["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"]) * SCALE[units["dram__bytes_write.sum"]]
                tag = short
                if short == "k_stats":
                    tag = "k_stats" if rep.startswith("r2_stats_full") else "k_stats_incremental"
                traffic[workload].setdefault(tag, int(rd + wr))
                lines.append("%s   [%s]" % (name, tag))
                for k in KEYS:
                    if k in d:
                        lines.append("  %-96s %s %s" % (k, d[k], units.get(k, "")))
                lines.append("  %-96s %d" % ("dram bytes read + written per launch", int(rd + wr)))
                lines.append("")
        with open(os.path.join(ROOT, "profiles", "r2_ncu_full_%s_summary.txt" % workload.lower()), "w") as f:
            f.write("\n".join(lines))
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
