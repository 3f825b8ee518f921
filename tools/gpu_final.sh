#!/bin/bash
# Final-tree cycle: parity tests, bench line + reference arm, every configuration at both resolutions, launch list, two --set full captures.
# Every step under its own timeout.  usage: tools/gpu_final.sh <tag>
tag=${1:-fin}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"
timeout 150 python tools/bench_configs.py 240 320 8192 > gpurun_out/configs_240x320_$tag.json 2> gpurun_out/configs_$tag.err; echo "configs240 rc=$?"
timeout 150 python tools/bench_configs.py 120 160 16384 > gpurun_out/configs_120x160_$tag.json 2>> gpurun_out/configs_$tag.err; echo "configs120 rc=$?"
A="--steps 2 --warmup 3 --quick"
timeout 120 python bench.py $A > gpurun_out/plain_$tag.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $A > gpurun_out/ncu_ll_$tag.log 2>&1; echo "launch list rc=$?"
A="--steps 1 --warmup 3 --quick"
timeout 120 python bench.py $A > gpurun_out/plain2_$tag.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py $A > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
A="--workload full_house_mask_240x320 --steps 1 --warmup 3 --frames 2048 --quick"
timeout 120 python bench.py $A > gpurun_out/plain3_$tag.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_preprocess -s 3 -c 1 -f -o gpurun_out/prof240_$tag python bench.py $A > gpurun_out/ncu240_$tag.log 2>&1; echo "ncu240 rc=$?"
