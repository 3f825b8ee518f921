#!/bin/bash
# ncu --set full capture of the fused kernel only.  usage: tools/gpu_ncu.sh <tag> [frames] [extra bench args]
tag=${1:-x}; nf=${2:-8192}; shift; shift
python bench.py --steps 1 --warmup 3 --frames $nf --quick "$@" > gpurun_out/plain2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --frames $nf --quick "$@" > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
