#!/bin/bash
# The driver's multi-GPU command for N ranks.  usage: tools/gpu_bench_n.sh <N> <tag>
n=$1; tag=$2
mkdir -p gpurun_out
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err; echo "bench rc=$? wall ${SECONDS}s"
tail -4 gpurun_out/bench_${n}gpu_$tag.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${n}gpu_$tag.json") if l.startswith("{")][0])
print("n_gpus", d["n_gpus"], "value", d["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("stats_on", {k: d.get("stats_on", {}).get(k) for k in ("ms_per_step", "delta_ms_vs_stats_off", "collective")})
e=d.get("e2e", {}); print("e2e", e.get("value"), "per rank", e.get("value_per_rank"), "ceiling/rank", e.get("pcie_ceiling_per_rank"), "tub", e.get("tub_mode", {}).get("value", e.get("tub_mode")))
for k, b in (d.get("blocks") or {}).items(): print(k, b["value"], b["ms_per_step"], b["roofline"]["frac"], (b.get("clocks") or {}).get("sm_mhz"))
PY
