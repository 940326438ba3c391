#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow" -p no:cacheprovider > gpurun_out/r2g_grow.log 2>&1; echo "grow rc=$?"; tail -3 gpurun_out/r2g_grow.log | cut -c1-600
for v in "" "GLB_SE_CL16=1"; do echo "== glue $v"; env $v timeout 300 python tools/glue_bw.py > gpurun_out/r2g_glue_$v.txt 2>&1; grep -i "style" gpurun_out/r2g_glue_$v.txt | head -10; done
GLB_SE_CL16=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "style_epilogue_vs_contract" -p no:cacheprovider 2>&1 | tail -2
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench.err
GLB_SE_CL16=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2g_bench_cl16.json 2> gpurun_out/r2g_bench_cl16.err; echo "bench cl16 rc=$?"
GLB_RESNET_GRAPHS=1 timeout 300 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg5_graphs.json 2> gpurun_out/r2g_bench_cfg5_graphs.err; echo "cfg5 graphs rc=$?"; tail -2 gpurun_out/r2g_bench_cfg5_graphs.err
timeout 300 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg5.json 2> gpurun_out/r2g_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
for f in ("r2g_bench","r2g_bench_cl16","r2g_bench_cfg5_graphs","r2g_bench_cfg5"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median")}, d["e2e"]["value"], round(d["roofline"]["achieved"],1), round(d["roofline_glue"]["achieved"],1), d.get("torch_eager_b200"), d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
