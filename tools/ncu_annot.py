#!/usr/bin/env python
"""Annotated SASS of one kernel from an .ncu-rep: executions per frame, stall samples, SASS text, CUDA source line.
usage: ncu_annot.py <report> <lib.so> <kernel-substring> <frames> > listing.txt"""
import csv, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

def main():
    rep, lib, ksub, frames = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    sl = ncu_lines.sass_lines(lib, ksub)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr = rows[1]; body = rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    assert len(sl) == len(body), (len(sl), len(body))
    ts = sum(int(r[ci['# Samples']] or 0) for r in body)
    for (addr, text, src), r in zip(sl, body):
        inst = int(r[ci['Instructions Executed']] or 0); smp = int(r[ci['# Samples']] or 0)
        s = f'{src[0]}:{src[1]}' if src else ''
        print(f'{addr:05x} {inst / frames:8.1f} {100.0 * smp / ts:5.2f}%  {text:70s} {s}')

if __name__ == '__main__':
    main()
