#!/bin/bash
# Profiles of the round: launch list of the default step + ncu --set full captures of the fused kernels (each only after the same command ran clean).
# usage: tools/gpu_prof.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
A="--steps 2 --warmup 3 --quick"
python bench.py $A > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $A > gpurun_out/ncu_ll_$tag.log 2>&1; echo "launch list rc=$?"
A="--steps 1 --warmup 3 --quick"
ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py $A > gpurun_out/ncu_$tag.log 2>&1; echo "ncu 65536 rc=$?"
A="--workload full_house_mask_240x320 --steps 1 --warmup 3 --frames 8192 --quick"
python bench.py $A > gpurun_out/plain3_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof240_$tag python bench.py $A > gpurun_out/ncu240_$tag.log 2>&1; echo "ncu240 rc=$?"
ls -la gpurun_out/*$tag*
