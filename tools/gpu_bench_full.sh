#!/bin/bash
# The driver's own command (full default bench) plus the reference arm.  usage: tools/gpu_bench_full.sh <tag> [extra args]
tag=${1:-full}; shift
mkdir -p gpurun_out
SECONDS=0; python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
echo "wall ${SECONDS}s"; tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_$tag.json") if l.startswith("{")][0])
print("value", d["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("stats_on", {k: d.get("stats_on", {}).get(k) for k in ("ms_per_step", "delta_ms_vs_stats_off")})
e=d.get("e2e", {}); print("e2e", e.get("value"), "ceiling/rank", e.get("pcie_ceiling_per_rank"), "tub", e.get("tub_mode", {}).get("value", e.get("tub_mode")))
for k, b in (d.get("blocks") or {}).items(): print(k, b["value"], b["ms_per_step"], b["roofline"]["frac"], (b.get("clocks") or {}).get("sm_mhz"))
print(json.dumps(d.get("other_workloads", {}).get("waypoint_speed_1M_states"), indent=None)[:1500])
PY
