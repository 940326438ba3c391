#!/usr/bin/env bash
# Round-2 evidence after the upsample / pool folds: smoke, full GPU suite, launch list, ncu --set full raw pages, sanitizer, benches
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py --smoke > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; grep "^\[smoke\]" $O/r2f_smoke.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2f_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2f_gpu_tests.log | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_cfg2_tf32.csv python tools/profile_step.py cfg2 tf32 > $O/r2f_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py $O/r2_launches_cfg2_tf32.csv > $O/r2_launches_cfg2_tf32_summary.txt 2>&1; head -12 $O/r2_launches_cfg2_tf32_summary.txt | cut -c1-150
full() {  # name regex count
  timeout 500 ncu --set full --clock-control none --import-source off --profile-from-start off -k "regex:$2" -c "$3" -f -o $O/full_$1 python tools/profile_step.py cfg2 tf32 > $O/r2f_full_$1.log 2>&1
  echo "full $1 rc=$?"
  ncu -i $O/full_$1.ncu-rep --page raw --csv > $O/r2_ncu_full_$1.csv 2>/dev/null
  rm -f $O/full_$1.ncu-rep
  python - "$1" <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/r2_ncu_full_{sys.argv[1]}.csv")))
h = rows[0]
def col(n): return h.index(n) if n in h else None
cols = {k: col(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed")}
for r in rows[2:]:
    print("   ", " | ".join(f"{r[i][:48]}" for k, i in cols.items() if i is not None))
PY
}
full conv_family "conv_(fprop|wgrad)_tc" 24
full folds_glue "fold_|blur_tile|act_bwd" 12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "contract" -p no:cacheprovider > $O/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r2_sanitizer_memcheck.log | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "upconv_family or downconv_family or blur_tile or folded" -p no:cacheprovider > $O/r2f_sanitizer_racecheck_folds.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/r2f_sanitizer_racecheck_folds.log | cut -c1-200
timeout 600 python bench.py > $O/r2_bench_cfg2_tf32.json 2> $O/r2f_bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 300 python bench.py --conv-impl bf16 --no-cpu-baseline > $O/r2_bench_cfg2_bf16.json 2> $O/r2f_bench_bf16.err; echo "bench bf16 rc=$?"
for c in cfg1 cfg3 cfg4 cfg5; do timeout 400 python bench.py --config $c --no-cpu-baseline > $O/r2_bench_$c.json 2> $O/r2f_bench_$c.err; echo "bench $c rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_cfg*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        r=d.get("roofline") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], round(r.get("achieved",0),1), round(r.get("frac",0),3), round((d.get("roofline_glue") or {}).get("achieved") or 0,1), (d.get("cpu_baseline") or {}).get("value"), (d.get("opt_in_bf16_operands") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
