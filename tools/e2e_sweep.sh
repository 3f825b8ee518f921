#!/bin/bash
for mb in 4 8 16 24 48 96; do echo "== chunk $mb MB"; TRS_HOST_CHUNK_MB=$mb python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-others 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['e2e']['value'])"; done
