#!/usr/bin/env bash
# one 8-GPU box: BASELINE.json configs[1..4] at their GPU counts
set -u
mkdir -p gpurun_out
run() { # cfg n extra...
  cfg=$1; n=$2; shift; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$((RANDOM % 9)) bench.py --config $cfg --gpus $n --steps 30 --warmup 5 > gpurun_out/r2k_${cfg}_n$n.json 2> gpurun_out/r2k_${cfg}_n$n.err; echo "$cfg n$n rc=$?"
}
run cfg2 8 A=1
run cfg3 8 A=1
run cfg4 8 A=1
run cfg5 8 A=1
run cfg3 4 A=1
run cfg3 2 A=1
run cfg2 4 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"])
    except Exception as e: print(f,"unreadable",e)
PY
