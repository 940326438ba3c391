#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "blur or glue_kernels" -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
for v in "A=1" "GLB_GLUE_OLD=4"; do echo "== glue $v"; env $v timeout 300 python tools/glue_bw.py > gpurun_out/r2q_glue_$v.txt 2>&1; grep -i "blur" gpurun_out/r2q_glue_$v.txt | head -10; done
for v in "A=1" "GLB_GLUE_OLD=4"; do env $v timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2q_bench_$v.json 2> gpurun_out/r2q_bench_$v.err; echo "bench $v rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q_bench_*.json")):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], round(d["roofline_glue"]["achieved"],1), {k:v for k,v in d["roofline_glue"]["by_kind_gbs"].items() if "blur" in k})
PY
