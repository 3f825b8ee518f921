#!/usr/bin/env python
"""Attribute instructions / stall samples of one kernel to the phase functions of preproc_fast.cuh (SASS order:
inlined helper lines inherit the phase of the last phase-function line seen).
usage: ncu_phases.py <report> <lib.so> <kernel-substring> <frames>"""
import csv, re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'triton-racer-sim_b200', 'csrc')
FILES = ('preproc_fast.cuh', 'preproc_bsw.cuh')


def phase_tables():
    return {f: phase_table(os.path.join(CSRC, f)) for f in FILES if os.path.exists(os.path.join(CSRC, f))}


def phase_of(tables, src):
    """src = (file, line) of a SASS instruction -> phase name or None"""
    if not src or src[0] not in tables:
        return None
    cur = None
    for ln, name in tables[src[0]]:
        if src[1] >= ln: cur = name
    return cur


def phase_table(src_path):
    names = {'p2_nms_lagged': 'p2', 'k_preprocess_bsw': 'main', 'wait_progress': 'sync', 'hsv_masks_of': 'hsv', 'p1_strip_walk': 'p1', 'p1b_colour_masks': 'p1b', 'p2_nms': 'p2', 'p3_hysteresis': 'p3', 'p3_relax_band': 'p3', 'p4_output': 'p4',
             'init_tables': 'prolog', 'k_preprocess_fast': 'main', 'k_preprocess_sw': 'main', 'k_preprocess_banded': 'main'}
    marks = []
    for i, ln in enumerate(open(src_path), 1):
        m = re.match(r'(__device__ __forceinline__ \w+ |__global__ void __launch_bounds__\([^)]*\) |__global__ void __maxnreg__\([^)]*\) )(\w+)\(', ln)
        if m and m.group(2) in names:
            marks.append((i, names[m.group(2)]))
    return marks

def main():
    rep, lib, ksub, frames = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    tables = phase_tables()
    sl = ncu_lines.sass_lines(lib, ksub)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr = rows[1]; body = rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    assert len(sl) == len(body), (len(sl), len(body))
    cur = '?'; agg = {}; ti = ts = 0
    for (addr, text, src), r in zip(sl, body):
        inst = int(r[ci['Instructions Executed']] or 0); smp = int(r[ci['# Samples']] or 0)
        ph = phase_of(tables, src)
        if ph: cur = ph
        a = agg.setdefault(cur, [0, 0]); a[0] += inst; a[1] += smp; ti += inst; ts += smp
    print(f'total warp-instr per frame {ti / frames:.0f}')
    for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'{k:8s} instr/frame {i / frames:9.0f} ({100 * i / ti:4.1f}%)  samples {100 * s / ts:4.1f}%')

if __name__ == '__main__':
    main()
