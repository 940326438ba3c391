#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for c in cfg4 cfg3; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_${c}_tf32.csv python tools/profile_step.py $c tf32 > gpurun_out/r2n_ncu_$c.log 2>&1; echo "$c launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_${c}_tf32.csv > gpurun_out/r2_launches_${c}_tf32_summary.txt 2>&1; head -32 gpurun_out/r2_launches_${c}_tf32_summary.txt | cut -c1-150
done
