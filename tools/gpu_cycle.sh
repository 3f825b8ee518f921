#!/bin/bash
# One GPU cycle: parity tests, bench line, then ncu launch list + full capture of the hot kernel.  usage: tools/gpu_cycle.sh <tag> [frames-for-ncu]
tag=${1:-x}; nf=${2:-8192}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json
python bench.py --steps 2 --warmup 3 --frames $nf --no-cpu-baseline --no-e2e --no-others > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --frames $nf --no-cpu-baseline --no-e2e --no-others > gpurun_out/ncu_ll_$tag.log 2>&1
python bench.py --steps 1 --warmup 3 --frames $nf --no-cpu-baseline --no-e2e --no-others > gpurun_out/plain2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_preprocess_fast -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --frames $nf --no-cpu-baseline --no-e2e --no-others > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
