#!/bin/bash
# sweep an environment knob over the headline bench.  usage: tools/gpu_sweep.sh "<VAR1=a VAR2=b>" "<VAR1=c ...>" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-others 2>&1 | grep -o '"value": [0-9.]*' | head -1
done
