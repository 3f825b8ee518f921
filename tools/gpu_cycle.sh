#!/bin/bash
# One full GPU cycle: parity tests, the bench line, then the ncu launch list and --set full captures of the fused kernels.
# usage: tools/gpu_cycle.sh <tag> [frames-for-ncu]
tag=${1:-x}; nf=${2:-8192}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"
A="--steps 2 --warmup 3 --frames $nf --quick"
python bench.py $A > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $A > gpurun_out/ncu_ll_$tag.log 2>&1
A="--steps 1 --warmup 3 --frames $nf --quick"
python bench.py $A > gpurun_out/plain2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py $A > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
A="--workload full_house_mask_240x320 --steps 1 --warmup 3 --frames 2048 --quick"
python bench.py $A > gpurun_out/plain3_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof240_$tag python bench.py $A > gpurun_out/ncu240_$tag.log 2>&1; echo "ncu240 rc=$?"
tools/ubench_store > gpurun_out/ubench_store_$tag.txt 2>&1
