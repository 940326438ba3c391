#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/parity
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "style or instance or rgb" -p no:cacheprovider > gpurun_out/r2f_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -6 gpurun_out/r2f_kernels.log | cut -c1-300
for m in 0 1; do GLB_SE_MODE=$m timeout 300 python tools/glue_bw.py > gpurun_out/r2f_glue_mode$m.txt 2>&1; echo "== SE mode $m"; grep -i "style\|rgb" gpurun_out/r2f_glue_mode$m.txt | head -24; done
GLB_DUMP_PARITY=gpurun_out/parity timeout 600 python -m pytest tests/test_cfg2_fullwidth.py -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2f_cfg2.log 2>&1; echo "cfg2 rc=$?"
grep -n "^E  \|passed\|failed" gpurun_out/r2f_cfg2.log | cut -c1-300 | head -12
timeout 600 python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow" -p no:cacheprovider > gpurun_out/r2f_grow.log 2>&1; echo "grow rc=$?"; tail -3 gpurun_out/r2f_grow.log | cut -c1-300
for m in 0 1 2; do GLB_SE_MODE=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_se$m.json 2> gpurun_out/r2f_bench_se$m.err; echo "bench se$m rc=$?"; done
python - <<'PY'
import json
for m in (0,1,2):
    d=json.load(open(f"gpurun_out/r2f_bench_se{m}.json"))
    print(m, {k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], round(d["roofline"]["achieved"],1), round(d["roofline_glue"]["achieved"],1), {k:v for k,v in d["roofline_glue"]["by_kind_gbs"].items() if "style" in k or "rgb" in k})
PY
