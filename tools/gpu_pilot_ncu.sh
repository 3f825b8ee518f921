#!/bin/bash
# Per-kernel times of the pilot forward pass (launch list) for both conv2 / conv3 formulations + a full capture of the row kernel.
tag=${1:-pl}
mkdir -p gpurun_out
for m in 1 0; do
  TRS_PILOT_ROWCONV=$m ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pilot -c 60 --csv --log-file gpurun_out/pilot_launches_${m}_$tag.csv python tools/pilot_bench.py 8192 8192 > gpurun_out/pilot_ncu_${m}_$tag.log 2>&1; echo "ncu($m) rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:k_pilot_rowconv -s 4 -c 2 -f -o gpurun_out/prof_rowconv_$tag python tools/pilot_bench.py 8192 8192 > gpurun_out/pilot_ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import csv, collections
for m in (1, 0):
    rows = [r for r in csv.reader(open(f"gpurun_out/pilot_launches_{m}_$tag.csv")) if len(r) > 10 and r[0].isdigit()]
    # last forward pass: the last 10 kernels
    print("ROWCONV", m)
    for r in rows[-10:]:
        print("  ", r[4][:60], r[7], r[8], r[-1], r[-2])
PY
