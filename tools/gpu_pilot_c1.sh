#!/bin/bash
tag=${1:-c1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pilot_conv1 -s 2 -c 1 -f -o gpurun_out/prof_conv1_$tag python tools/pilot_bench.py 8192 8192 > gpurun_out/pilot_ncu_c1_$tag.log 2>&1; echo "ncu rc=$?"
