#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/parity
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "style or instance or rgb or conv" -p no:cacheprovider > gpurun_out/r2e_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -6 gpurun_out/r2e_kernels.log | cut -c1-300
for m in 0 1 2; do GLB_SE_MODE=$m timeout 300 python tools/glue_bw.py > gpurun_out/r2e_glue_mode$m.txt 2>&1; echo "== SE mode $m"; grep -i "style\|rgb" gpurun_out/r2e_glue_mode$m.txt | head -24; done
GLB_DUMP_PARITY=gpurun_out/parity timeout 600 python -m pytest tests/test_cfg2_fullwidth.py -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2e_cfg2.log 2>&1; echo "cfg2 rc=$?"
grep -n "^E  \|passed\|failed" gpurun_out/r2e_cfg2.log | cut -c1-300 | head -12
timeout 600 python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow" -p no:cacheprovider > gpurun_out/r2e_grow.log 2>&1; echo "grow rc=$?"; tail -3 gpurun_out/r2e_grow.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench.json"))
print({k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline_glue"]["achieved"], d["roofline_glue"]["by_kind_gbs"])
PY
