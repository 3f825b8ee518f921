#!/bin/bash
# One optimisation iteration on the GPU: parity tests, the headline bench line, then the warp-instruction count of the fused kernel.
# usage: tools/gpu_iter.sh <tag> [pytest -k expression]
tag=${1:-it}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_$tag.log 2>&1; else python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; fi
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 3 --quick > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench_$tag.json; tail -2 gpurun_out/bench_$tag.err
A="--steps 1 --warmup 3 --frames 8192 --quick"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_preprocess -s 3 -c 1 --csv --log-file gpurun_out/inst_$tag.csv python bench.py $A > gpurun_out/ncu_inst_$tag.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/inst_$tag.csv
