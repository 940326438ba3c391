#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$((RANDOM % 9)) bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/r2p_${tag}_n$n.json 2> gpurun_out/r2p_${tag}_n$n.err; echo "$tag n$n rc=$?"; }
run overlap 8 GLB_DP_OVERLAP_D=1
run serial 8 GLB_DP_OVERLAP_D=0
run overlap 4 GLB_DP_OVERLAP_D=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median")}, d["e2e"]["value"])
    except Exception as e: print(f,"unreadable",e)
PY
