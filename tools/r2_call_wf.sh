#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "folded or downconv or upconv" -p no:cacheprovider 2>&1 | tail -8 | cut -c1-400
timeout 300 python tools/check_upconv.py 2>&1 | grep "wgrad" | tail -6
for v in "A=1" "GLB_WGRAD_FOLD2=0"; do env $v timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; print('$v', d['value'], d['ms_per_step'], r['conv_ms_per_step'], round(r['achieved'],1), {k:round(v) for k,v in r['by_kind_tflops'].items() if 'wgrad' in k})"; done
