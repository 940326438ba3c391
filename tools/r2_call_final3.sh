#!/usr/bin/env bash
# final evidence refresh after the fold tap-reuse kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py --smoke > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; grep "^\[smoke\]" $O/r2f_smoke.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2f_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2f_gpu_tests.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "upconv_family or downconv_family or folded" -p no:cacheprovider > $O/r2_sanitizer_memcheck_folds.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r2_sanitizer_memcheck_folds.log | cut -c1-200
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_cfg2_tf32.csv python tools/profile_step.py cfg2 tf32 > $O/r2f_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py $O/r2_launches_cfg2_tf32.csv > $O/r2_launches_cfg2_tf32_summary.txt 2>&1; head -10 $O/r2_launches_cfg2_tf32_summary.txt | cut -c1-150
timeout 300 python tools/check_upconv.py > $O/r2_check_upconv.txt 2>&1
timeout 600 python bench.py > $O/r2_bench_cfg2_tf32.json 2> $O/r2f_bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 300 python bench.py --conv-impl bf16 --no-cpu-baseline > $O/r2_bench_cfg2_bf16.json 2> $O/r2f_bench_bf16.err; echo "bench bf16 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cfg2_tf32.json","gpurun_out/r2_bench_cfg2_bf16.json"):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]
    print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","gpu_launches")}, d["e2e"]["value"], round(r["achieved"],1), round(r["frac"],3), round(r["tflops_on_literal_sequence_flops"],1), r["conv_ms_per_step"], round(d["roofline_glue"]["achieved"],1), round(d["roofline_glue"]["frac"],3), (d.get("cpu_baseline") or {}).get("value"), (d.get("opt_in_bf16_operands") or {}).get("value"), (d.get("torch_eager_b200") or {}).get("value"))
PY
