#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run2() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM % 9)) bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2o_n2_$tag.json 2> gpurun_out/r2o_n2_$tag.err; echo "n2 $tag rc=$?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2o_n2_$tag.err | tail -3 | cut -c1-300; }
run2 overlap GLB_DP_OVERLAP_D=1
run2 serial GLB_DP_OVERLAP_D=0
run2 overlap_nograph GLB_DP_OVERLAP_D=1 GLB_X=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_n2_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median")}, d["e2e"]["value"], d["details"]["grad_allreduce"][-60:])
    except Exception as e: print(f,"unreadable",e)
PY
