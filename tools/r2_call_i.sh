#!/usr/bin/env bash
# 2 GPUs: data-parallel A/B (in-place grouped all-reduce vs packed buckets), other configs at 1 GPU
set -u
mkdir -p gpurun_out
run2() { # tag env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2i_n2_$tag.json 2> gpurun_out/r2i_n2_$tag.err; echo "n2 $tag rc=$?"; tail -2 gpurun_out/r2i_n2_$tag.err | cut -c1-300
}
run2 inplace GLB_DP_INPLACE=1
run2 packed GLB_DP_INPLACE=0
run2 inplace_b128 GLB_DP_INPLACE=1 GLB_DP_BUCKET_MB=128
for c in cfg1 cfg3 cfg4; do timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_$c.json 2> gpurun_out/r2i_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r2i_$c.err | cut -c1-300; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_*.json")):
    try:
        d=json.load(open(f)); print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], (d.get("roofline") or {}).get("achieved"), (d.get("roofline_glue") or {}).get("achieved"))
    except Exception as e: print(f,"unreadable",e)
PY
