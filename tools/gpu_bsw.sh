#!/bin/bash
# 240x320 store-warp banded kernel: parity tests first (under a timeout: the kernel spins on progress words), then the two 240x320 workloads.
# usage: tools/gpu_bsw.sh <tag> [frames]
tag=${1:-bsw}; nf=${2:-16384}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fresh_frames or golden or full_size or other_kernels or statistics" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
for wl in full_house_mask_240x320 full_chain_240x320; do
  timeout 300 python bench.py --workload $wl --quick --steps 10 --warmup 3 --frames $nf > gpurun_out/bench_${wl}_$tag.json 2> gpurun_out/bench_${wl}_$tag.err; echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${wl}_$tag.json") if l.startswith("{")][0])
    print("$wl", round(d["value"]/1e6,3), "M frames/s  frac", round(d["roofline"]["frac"],4), "ms", round(d["ms_per_step"],3))
except Exception as e: print("no line", e)
PY
done
