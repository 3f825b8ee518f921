#!/bin/bash
# Quick GPU check: parity tests + the headline bench line only.  usage: tools/gpu_quick.sh <tag> [extra bench args]
tag=${1:-q}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 --quick "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
