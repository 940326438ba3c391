#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cat > /tmp/in_one.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gan_lab_b200 import _kernels as K
hs, ho, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
src = torch.randint(0, 256, (n, hs, hs, 3), dtype=torch.uint8, device="cuda")
idx = torch.randperm(n)
for _ in range(3):
    out = K.u8_box_resize_normalize(src, idx, (ho, ho), (.5,) * 3, (.5,) * 3)
torch.cuda.synchronize()
PY
for cfg in "1024 128 96" "128 128 2048"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source off -k "regex:grouped" -s 2 -c 1 -f -o gpurun_out/in_$1_$2 python /tmp/in_one.py $1 $2 $3 > gpurun_out/r2in2_$1_$2.log 2>&1; echo "ncu $1->$2 rc=$?"
  ncu -i gpurun_out/in_$1_$2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_input_$1_$2.csv 2>/dev/null
  rm -f gpurun_out/in_$1_$2.ncu-rep
done
