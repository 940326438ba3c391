#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "folded or downconv or upconv or bf16" -p no:cacheprovider 2>&1 | tail -8 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | cut -c1-400
for v in tf32 bf16; do timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --conv-impl $v > gpurun_out/r2z_bench_$v.json 2> gpurun_out/r2z_bench_$v.err; echo "bench $v rc=$?"; done
GLB_BATCH_D=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2z_bench_batchd.json 2> gpurun_out/r2z_bench_batchd.err; echo "bench batchd rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_bench*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        r=d["roofline"]
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], round(r["achieved"],1), round(r["frac"],3), round(r.get("tflops_on_literal_sequence_flops",0),1), r.get("conv_ms_per_step"), round(d["roofline_glue"]["achieved"],1), d["roofline_glue"]["glue_ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
