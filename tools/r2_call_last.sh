#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep "^\[smoke\]" | cut -c1-160
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"
