#!/usr/bin/env bash
# after the upsample / pool folds: op test, launch list of one cfg2 step, per-shape conv dump, per-layer timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "downconv_op or upconv_op" -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_cfg2_tf32.csv python tools/profile_step.py cfg2 tf32 > gpurun_out/r2u_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2u_launches_cfg2_tf32.csv > gpurun_out/r2u_launches_cfg2_tf32_summary.txt 2>&1; head -45 gpurun_out/r2u_launches_cfg2_tf32_summary.txt | cut -c1-160
GLB_BENCH_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_dump.json 2> gpurun_out/r2u_bench_dump.err; grep "^\[dump\]" gpurun_out/r2u_bench_dump.err | cut -c1-220
