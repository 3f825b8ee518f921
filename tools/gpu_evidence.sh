#!/bin/bash
# Evidence run: every golden configuration at 240x320, the tub-ingestion kernels under ncu.  usage: tools/gpu_evidence.sh <tag>
tag=${1:-ev}
mkdir -p gpurun_out
python tools/bench_configs.py 240 320 8192 > gpurun_out/configs_240x320_$tag.json 2> gpurun_out/configs_$tag.err; echo "configs rc=$?"
python tools/jpeg_bench.py 65536 > gpurun_out/jpeg_bench_$tag.log 2>&1; echo "jpeg rc=$?"; cat gpurun_out/jpeg_bench_$tag.log
ncu --set full --clock-control none --import-source on -k regex:k_jpeg -s 6 -c 3 -f -o gpurun_out/prof_jpeg_$tag python tools/jpeg_bench.py 65536 > gpurun_out/ncu_jpeg_$tag.log 2>&1; echo "ncu rc=$?"
