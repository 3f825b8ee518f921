#!/usr/bin/env python
"""Markdown summary of one kernel of an .ncu-rep for profiles/: duration, DRAM traffic, pipe utilisation, stall reasons and
instructions per frame by phase.  usage: ncu_report.py <report.ncu-rep> <lib.so> <kernel-substring> <frames> [title]"""
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def main():
    rep, lib, ksub, frames = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    title = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep)
    d, u = raw(rep)
    f = lambda k: float(d[k]) if d.get(k) not in (None, '') else float('nan')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    rd = f('dram__bytes_read.sum') * scale[u['dram__bytes_read.sum']]
    wr = f('dram__bytes_write.sum') * scale[u['dram__bytes_write.sum']]
    dur_us = f('gpu__time_duration.sum') * {'us': 1.0, 'ms': 1e3, 'ns': 1e-3}[u['gpu__time_duration.sum']]
    print(f"# {title}\n")
    print(f"`ncu --set full --clock-control none`, one launch of `{d.get('Kernel Name', ksub)}` over {int(frames)} frames "
          f"(grid {d.get('launch__grid_size')}, block {d.get('launch__block_size')}, {d.get('launch__registers_per_thread')} registers, "
          f"{d.get('launch__shared_mem_per_block_dynamic')} {u.get('launch__shared_mem_per_block_dynamic')} dynamic shared memory).  "
          "Times under the profiler are cold-cache and not bench values.\n")
    print("| metric | value |\n|---|---|")
    print(f"| duration | {dur_us:.1f} us ({frames / dur_us:.2f} M frames/s under ncu) |")
    print(f"| DRAM read / write | {rd / 1e6:.1f} MB / {wr / 1e6:.1f} MB = {(rd + wr) / frames:.0f} B per frame |")
    for k, name in [('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput (% of peak)'),
                    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput (% of peak)'),
                    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy (%)'),
                    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe (%)'),
                    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe (%)'),
                    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe (%)'),
                    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory wavefronts (% of peak)'),
                    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy (%)'),
                    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe (%)')]:
        if k in d:
            print(f"| {name} | {f(k):.1f} |")
    print(f"| warp instructions per frame | {f('smsp__inst_executed.sum') / frames:.0f} |")
    print()
    import ncu_stalls
    import ncu_phases
    sys.argv = ['ncu_phases.py', rep, lib, ksub, str(frames)]
    print("## Warp instructions and stall samples by phase\n\n```")
    ncu_phases.main()
    print("```\n\n## Stall reasons by phase (warp-state samples)\n\n```")
    sys.argv = ['ncu_stalls.py', rep, lib, ksub]
    ncu_stalls.main()
    print("```")


if __name__ == '__main__':
    main()
