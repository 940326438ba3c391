#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py > gpurun_out/r2dp4_check.log 2>&1; echo "dp_check rc=$?"
grep -v "^W0\|^\[W\|NCCL version" gpurun_out/r2dp4_check.log | grep -B2 -A12 "Traceback\|dp_check world\|Error" | head -60
