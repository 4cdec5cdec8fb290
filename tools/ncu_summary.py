#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline discussion needs."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "lim_reg"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wave%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "st_short"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "st_bar"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "st_lg"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "st_mio"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "st_math"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "st_wait"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "st_notsel"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("###", r[kn][:90])
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                parts.append(f"{short}={r[i]}{units[i] if units[i] not in ('', '%') else ''}")
        print("   " + "  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
