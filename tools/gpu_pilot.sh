#!/bin/bash
# Pilot networks: parity tests (both formulations of conv2 / conv3), then throughput with each.  usage: tools/gpu_pilot.sh <tag>
tag=${1:-pl}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pilot_gpu.py -x -q -m gpu > gpurun_out/pytest_pilot_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_pilot_$tag.log
for m in 1 0; do
  echo "TRS_PILOT_ROWCONV=$m"
  TRS_PILOT_ROWCONV=$m timeout 300 python tools/pilot_bench.py 16384 8192 2>&1 | tail -2
done
