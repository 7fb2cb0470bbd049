#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the table profiles/ keeps: one row per captured launch with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.csv"""
import csv
import io
import subprocess
import sys

KEEP = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "l1_data_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1_data_wavefronts_pct"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"), ("smsp__issue_active.avg.per_cycle_active", "issue_per_cycle"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__inst_executed.sum", "instructions"),
]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow([name + (f" [{units[idx[m]]}]" if m in idx and units[idx[m]] else "") for m, name in KEEP])
    for r in data:
        w.writerow([r[idx[m]] if m in idx else "" for m, _ in KEEP])


if __name__ == "__main__":
    main()
