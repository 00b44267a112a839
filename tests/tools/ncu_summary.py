#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep, as text for profiles/.

    python tests/tools/ncu_summary.py gpurun_out/prof.ncu-rep "header line" > profiles/rNN_ncu_....txt
"""
import csv
import subprocess
import sys

WANT = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "launch__cluster_size", "launch__occupancy_cluster_pct", "launch__occupancy_cluster_gpu_pct",
    "launch__cluster_max_active", "launch__occupancy_limit_blocks",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    for line in sys.argv[2:]:
        print(line)
    print()
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    for row in rows[2:]:
        print("%-88s %s " % ("Kernel Name", row[ki]))
        for c in WANT:
            if c in head:
                i = head.index(c)
                print("%-88s %s %s" % (c, row[i], units[i]))
        print()


if __name__ == "__main__":
    main()
