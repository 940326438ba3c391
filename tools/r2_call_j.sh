#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "style_epilogue_vs_contract" -p no:cacheprovider 2>&1 | tail -2
for m in 0 1 3; do GLB_SE_MODE=$m timeout 300 python tools/glue_bw.py > gpurun_out/r2j_glue_mode$m.txt 2>&1; echo "== SE mode $m"; grep -i "style" gpurun_out/r2j_glue_mode$m.txt | head -8; done
for m in 1 3; do GLB_SE_MODE=$m timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_se$m.json 2> gpurun_out/r2j_bench_se$m.err; echo "bench se$m rc=$?"; done
python - <<'PY'
import json
for m in (1,3):
    d=json.loads([l for l in open(f"gpurun_out/r2j_bench_se{m}.json").read().splitlines() if l.startswith("{")][-1])
    print(m, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median")}, d["e2e"]["value"], round(d["roofline_glue"]["achieved"],1), {k:v for k,v in d["roofline_glue"]["by_kind_gbs"].items() if "style" in k})
PY
