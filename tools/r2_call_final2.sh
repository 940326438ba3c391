#!/usr/bin/env bash
# final evidence refresh: smoke, full GPU suite, launch list, headline bench line (with cpu baseline + bf16 sub-run)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py --smoke > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; grep "^\[smoke\]" $O/r2f_smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2f_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2f_gpu_tests.log | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_cfg2_tf32.csv python tools/profile_step.py cfg2 tf32 > $O/r2f_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py $O/r2_launches_cfg2_tf32.csv > $O/r2_launches_cfg2_tf32_summary.txt 2>&1; head -8 $O/r2_launches_cfg2_tf32_summary.txt | cut -c1-150
timeout 600 python bench.py > $O/r2_bench_cfg2_tf32.json 2> $O/r2f_bench_cfg2.err; echo "bench cfg2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_cfg2_tf32.json").read().splitlines() if l.startswith("{")][-1])
r=d["roofline"]
print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","gpu_launches")}, d["e2e"], round(r["achieved"],1), round(r["frac"],3), round(r["tflops_on_literal_sequence_flops"],1), round(d["roofline_glue"]["achieved"],1), round(d["roofline_glue"]["frac"],3), d["cpu_baseline"]["value"], d["opt_in_bf16_operands"]["value"], d["torch_eager_b200"]["value"], d["roofline_input"]["achieved"], d["clocks"])
PY
