#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_launches_cfg2_tf32.csv python tools/profile_step.py cfg2 tf32 > gpurun_out/r2x_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2x_launches_cfg2_tf32.csv > gpurun_out/r2x_launches_cfg2_tf32_summary.txt 2>&1; head -30 gpurun_out/r2x_launches_cfg2_tf32_summary.txt | cut -c1-160
for i in 1 2; do timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'])"; done
