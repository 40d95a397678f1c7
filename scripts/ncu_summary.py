#!/usr/bin/env python
"""Summarise an .ncu-rep — or the CSV of its raw page (`ncu -i rep --page raw --csv`) — into a small markdown table for
profiles/. Usage: ncu_summary.py rep|csv [title]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts % of peak"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global-load wavefronts"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "local-load wavefronts"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "local-store wavefronts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "FMA-heavy pipe %"),
    ("smsp__inst_executed_pipe_fp64.sum", "FP64 instructions"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    print(f"### {title}\n")
    print("| metric | unit | " + " | ".join(f"{r[ki].split('(')[0].replace('void ', '')} #{i}" for i, r in enumerate(data)) + " |")
    print("|---|---|" + "---|" * len(data))
    for key, label in KEYS:
        if key in H:
            j = H.index(key)
            vals = []
            for r in data:
                try:
                    vals.append(f"{float(r[j].replace(',', '')):.4g}")
                except ValueError:
                    vals.append(r[j])
            print(f"| {label} (`{key}`) | {units[j]} | " + " | ".join(vals) + " |")
    print()


if __name__ == "__main__":
    main()
