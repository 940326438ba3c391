#!/usr/bin/env bash
# upsample folded into the convolution: kernel contracts, op test, the conv family it shares kernels with, A/B bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "upconv or conv_family or blur or tf32" -p no:cacheprovider 2>&1 | tail -12 | cut -c1-400
for v in "GLB_UPCONV=1" "GLB_UPCONV=0"; do env $v timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err; echo "bench $v rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_bench_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"].get("conv_ms_per_step"), round(d["roofline_glue"]["achieved"],1), d["roofline_glue"]["glue_ms_per_step"], {k:v for k,v in d["roofline_glue"]["by_kind_gbs"].items() if "blur" in k or "upsample" in k})
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
