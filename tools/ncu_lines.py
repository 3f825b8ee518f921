#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals for one kernel of an .ncu-rep.

ncu's CSV source page is SASS-level; this joins it with `nvdisasm -g` line info from the cubin inside the .so
(instruction order is identical), then sums 'Instructions Executed' and stall samples per CUDA source line.
usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-name-substring> [top N]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    lines, cur, active = [], None, False
    for ln in out.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            active = kernel_sub in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return lines


def main():
    rep, lib, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    sl = sass_lines(lib, ksub)
    assert len(sl) == len(body), (len(sl), len(body))
    per = {}
    tot_inst = tot_samp = 0
    for (addr, text, src), r in zip(sl, body):
        inst = int(r[ci['Instructions Executed']] or 0)
        samp = int(r[ci['# Samples']] or 0)
        tot_inst += inst
        tot_samp += samp
        d = per.setdefault(src, [0, 0, {}])
        d[0] += inst
        d[1] += samp
        op = text.split()[0] if not text.startswith('@') else text.split()[1]
        op = op.split('.')[0]
        d[2][op] = d[2].get(op, 0) + inst
    print(f"total warp-instructions {tot_inst}, samples {tot_samp}")
    for src, (inst, samp, ops) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        topops = ' '.join(f"{k}:{v * 100 // max(inst, 1)}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
        print(f"{str(src):38s} inst {inst:>11d} {inst * 100.0 / tot_inst:5.1f}%  samples {samp * 100.0 / max(tot_samp, 1):5.1f}%  {topops}")
    # opcode histogram
    hist = {}
    for (addr, text, src), r in zip(sl, body):
        inst = int(r[ci['Instructions Executed']] or 0)
        op = text.split()[0] if not text.startswith('@') else text.split()[1]
        op = op.split('.')[0]
        hist[op] = hist.get(op, 0) + inst
    print("opcodes:", ' '.join(f"{k}:{v * 100.0 / tot_inst:.1f}%" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:24]))


if __name__ == '__main__':
    main()
