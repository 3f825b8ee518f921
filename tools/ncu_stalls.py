#!/usr/bin/env python
"""Per-phase stall-reason breakdown (warp-state samples) of one kernel.  usage: ncu_stalls.py <report> <lib.so> <kernel-substring>"""
import csv, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines, ncu_phases

def main():
    rep, lib, ksub = sys.argv[1:4]
    tables = ncu_phases.phase_tables()
    sl = ncu_lines.sass_lines(lib, ksub)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr = rows[1]; body = rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and not h.endswith('(Not Issued)')]
    assert len(sl) == len(body)
    cur = '?'; agg = {}; tot = 0
    for (addr, text, src), r in zip(sl, body):
        ph = ncu_phases.phase_of(tables, src)
        if ph: cur = ph
        a = agg.setdefault(cur, {})
        for c in stall_cols:
            v = int(r[ci[c]] or 0)
            a[c] = a.get(c, 0) + v; tot += v
    for ph, a in sorted(agg.items(), key=lambda kv: -sum(kv[1].values())):
        s = sum(a.values())
        top = ' '.join(f"{k[6:]}:{100.0 * v / s:.0f}%" for k, v in sorted(a.items(), key=lambda kv: -kv[1])[:7] if v)
        print(f'{ph:7s} {100.0 * s / tot:5.1f}% of samples | {top}')

if __name__ == '__main__':
    main()
