#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "downconv or upconv" -p no:cacheprovider 2>&1 | tail -5 | cut -c1-400
timeout 300 python tools/check_upconv.py > gpurun_out/r2w_check_upconv.txt 2>&1; tail -22 gpurun_out/r2w_check_upconv.txt
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2w_bench*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], round(d["roofline"]["achieved"],1), d["roofline"].get("conv_ms_per_step"), {k:round(v) for k,v in d["roofline"]["by_kind_tflops"].items()}, round(d["roofline_glue"]["achieved"],1), d["roofline_glue"]["glue_ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
