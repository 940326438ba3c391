#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py 2>&1 | grep -v "^W\|NCCL version" | tail -5
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$((RANDOM % 9)) bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/r2dp3_${tag}_n$N.json 2> gpurun_out/r2dp3_${tag}_n$N.err; echo "$tag n$N rc=$?"; }
run pack A=1
run nopack GLB_DP_PACK_BELOW=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2dp3_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","n_gpus")}, d["e2e"]["value"])
    except Exception as e: print(f,"unreadable",e); print(open(f.replace(".json",".err")).read()[-800:])
PY
