#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'sm__inst_executed.sum', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct', 'inst_executed_pipe',
        'pipe_alu_cycles_active', 'pipe_fma_cycles_active', 'pipe_fmaheavy', 'pipe_fmalite', 'bank_conflicts', 'warp_issue_stalled', 'smsp__inst_executed.avg.per_cycle_active',
        'launch__grid_size', 'launch__block_size', 'dynamic_smem', 'lsu_mem_shared', 'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'smsp__cycles_active.avg', 'achieved_occupancy', 'sm__warps_active']


def main(path, kernel_index=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    vals = rows[2 + kernel_index]
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS):
            v = vals[i]
            if 'warp_issue_stalled' in h and not h.endswith('_per_warp_active.pct'):
                continue
            print(f"{h} [{units[i]}] = {v}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
