#!/usr/bin/env bash
# data-parallel check after the folds: cfg2 on 2 GPUs (default switches) and with D's update overlapped
set -u
mkdir -p gpurun_out
N=${1:-2}
run() { tag=$1; cfg=$2; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$((RANDOM % 9)) bench.py --gpus $N --config $cfg --steps 40 --warmup 8 > gpurun_out/r2dp_${tag}_n$N.json 2> gpurun_out/r2dp_${tag}_n$N.err; echo "$tag n$N rc=$?"; }
run cfg2 cfg2 A=1
if [ "$N" = "2" ]; then run cfg2_overlap cfg2 GLB_DP_OVERLAP_D=1; fi
if [ "$N" = "8" ]; then run cfg3 cfg3 A=1; run cfg5 cfg5 A=1; fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2dp_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","n_gpus")}, d["e2e"]["value"])
    except Exception as e: print(f,"unreadable",e); print(open(f.replace(".json",".err")).read()[-800:])
PY
