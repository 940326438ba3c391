#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/parity
GLB_CHECK_IMPL=bf16 timeout 300 python tools/check_tc.py fprop dgrad wgrad > gpurun_out/r2h_tc_bf16.txt 2>&1; echo "check bf16 rc=$?"; cat gpurun_out/r2h_tc_bf16.txt | cut -c1-120
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bf16" -p no:cacheprovider > gpurun_out/r2h_bf16_tests.log 2>&1; echo "bf16 tests rc=$?"; tail -15 gpurun_out/r2h_bf16_tests.log | cut -c1-250
GLB_DUMP_PARITY=gpurun_out/parity timeout 600 python -m pytest tests/test_cfg2_fullwidth.py -m gpu -q -k "bf16" -rxXfE -p no:cacheprovider > gpurun_out/r2h_cfg2.log 2>&1; echo "cfg2 bf16 rc=$?"
grep -n "^E  \|passed\|failed" gpurun_out/r2h_cfg2.log | cut -c1-400 | head -8
timeout 300 python bench.py --conv-impl bf16 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_bf16.json 2> gpurun_out/r2h_bench_bf16.err; echo "bench bf16 rc=$?"; tail -3 gpurun_out/r2h_bench_bf16.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2h_bench_bf16.json"))
    print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","dtype")}, d["e2e"]["value"], d["roofline"], d["roofline_glue"]["achieved"], d["roofline_glue"]["by_kind_ms"])
except Exception as e: print("unreadable", e)
PY
