#!/usr/bin/env bash
# evidence: launch list of one cfg2 step, ncu --set full raw pages of the conv / style / other glue kernels, compute-sanitizer
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow" -p no:cacheprovider > gpurun_out/r2g2_grow.log 2>&1; echo "grow rc=$?"; tail -2 gpurun_out/r2g2_grow.log | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg2_tf32.csv python tools/profile_step.py cfg2 tf32 > gpurun_out/r2g2_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_cfg2_tf32.csv > gpurun_out/r2_launches_cfg2_tf32_summary.txt 2>&1; head -30 gpurun_out/r2_launches_cfg2_tf32_summary.txt
full() {  # name regex count
  timeout 500 ncu --set full --clock-control none --import-source off --profile-from-start off -k "regex:$2" -c "$3" -f -o gpurun_out/full_$1 python tools/profile_step.py cfg2 tf32 > gpurun_out/r2g2_full_$1.log 2>&1
  echo "full $1 rc=$?"
  ncu -i gpurun_out/full_$1.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$1.csv 2>/dev/null
  rm -f gpurun_out/full_$1.ncu-rep
  python - "$1" <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/r2_ncu_full_{sys.argv[1]}.csv")))
h = rows[0]
def col(n): return h.index(n) if n in h else None
cols = {k: col(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active")}
for r in rows[2:]:
    print("   ", " | ".join(f"{r[i][:60]}" for k, i in cols.items() if i is not None))
PY
}
full conv_dominant "conv_fprop_tc2_halo_kernel<256>" 2
full conv_family "conv_(fprop|wgrad)_tc" 16
full style "se_(fwd|bwd)_" 16
full glue "rgb_|blur3x3|act_bwd|pool|upsample|adam_ewma" 24
# compute-sanitizer: memcheck and racecheck over the per-kernel contract tests
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "contract" -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "style_epilogue_vs_contract or mbstd or blur or pixelnorm or rgb or conv_family_tf32" -p no:cacheprovider > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.log | cut -c1-200
