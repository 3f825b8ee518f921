#!/usr/bin/env python
"""profiles/ evidence: mnemonic census + trimmed SASS excerpts of the hot kernels from the in-tree library.
usage: tools/sass_excerpt.py <lib.so> > profiles/rNN_sass_excerpt.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
KERNELS = [("k_preprocess_swILi2ELi24ELi23ELi120ELi160", "k_preprocess_sw<2,24,23,120,160> (headline: fused chain, 120x160)",
            ["UBLKCP", "SYNCS", "HFMA2", "HSET2", "VIMNMX3", "BAR.ARV", "STG.E.128"]),
           ("k_preprocess_bandedILi2ELb1ELi24ELi23", "k_preprocess_banded<2,true,24,23> (240x320)", ["UBLKCP", "SYNCS", "HFMA2", "IDP.4A"]),
           ("k_pilot_gemmILi128ELi3ELb0", "k_pilot_gemm<128,3,false> (pilot convolutions, implicit GEMM)", ["UTMALDG", "UTCHMMA", "LDTM", "SYNCS", "UTCBAR"]),
           ("k_pilot_conv1", "k_pilot_conv1 (first convolution from u8 frames, A operand through tensor memory)", ["UTCHMMA", "STTM", "LDTM", "UTMALDG", "SYNCS"]),
           ("k_locate_grid", "k_locate_grid (nearest waypoint through the grid, fp64)", ["DADD", "DSETP", "LDS.128", "ATOMS", "ATOMG", "RED"]),
           ("k_locateEPKd", "k_locate (nearest waypoint, scanning kernel, fp64)", ["DADD", "DSETP", "LDS"]),
           ("k_preprocess_bswILi2ELi24ELi23ELi240ELi320ELi24ELi2ELi80ELb0", "k_preprocess_bsw<2,24,23,240,320,24,2,80,false> (240x320, default configuration)", ["UBLKCP", "SYNCS", "HFMA2", "HSET2", "VIMNMX3", "ATOMS"]),
           ("k_jpeg_entropy", "k_jpeg_entropy (tub ingestion)", ["SHF", "LDG", "IMAD"])]
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True)
arch = ", ".join(l.split(":", 1)[1].strip() for l in (arch.stdout + arch.stderr).splitlines() if ":" in l)
print("# SASS evidence, round 2\n")
print(f"`cuobjdump -sass {lib}` (built by `python -m triton_racer_sim_b200.build`: `-gencode arch=compute_100a,code=sm_100a -lineinfo`); "
      f"ELF images in the library: `{arch}` (sm_100a only).\n")
for pat, title, keys in KERNELS:
    out = subprocess.run(["bash", "tools/sass_of.sh", lib, pat], capture_output=True, text=True).stdout.splitlines()
    ops = collections.Counter()
    for l in out:
        t = l.split(" ", 1)[1]
        if t.startswith("@"):
            t = t.split(" ", 1)[1]
        ops[t.split()[0].rstrip(";")] += 1
    print(f"## {title}\n\n{len(out)} instructions.  Census of the mnemonics that identify the hardware paths used:\n")
    print("| mnemonic prefix | count |\n|---|---|")
    for k in keys:
        print(f"| `{k}` | {sum(v for o, v in ops.items() if o.startswith(k))} |")
    top = ", ".join(f"{o} {v}" for o, v in ops.most_common(14))
    print(f"\nMost frequent: {top}.\n")
    # first occurrence of the first key with some context
    for k in keys[:2]:
        idx = next((i for i, l in enumerate(out) if re.search(r"\b" + re.escape(k), l)), None)
        if idx is None:
            continue
        lo, hi = max(0, idx - 6), min(len(out), idx + 10)
        print(f"Excerpt around the first `{k}`:\n\n```\n" + "\n".join(out[lo:hi]) + "\n```\n")
