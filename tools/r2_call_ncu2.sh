#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --set full --clock-control none --import-source off --profile-from-start off -k "regex:conv_(fprop|wgrad)_tc" -c 40 -f -o $O/full_conv python tools/profile_step.py cfg2 tf32 > $O/r2ncu2.log 2>&1; echo "ncu rc=$?"
ncu -i $O/full_conv.ncu-rep --page raw --csv > $O/r2_ncu_full_conv_family.csv 2>/dev/null
rm -f $O/full_conv.ncu-rep
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_ncu_full_conv_family.csv")))
h = rows[0]
cols = [h.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed") if k in h]
for r in rows[2:]:
    print("   ", " | ".join(r[i][:52] for i in cols))
PY
