#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/parity
GLB_DUMP_PARITY=gpurun_out/parity python -m pytest tests/test_cfg2_fullwidth.py -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2b_cfg2.log 2>&1; echo "cfg2 rc=$?"
tail -30 gpurun_out/r2b_cfg2.log | cut -c1-600
python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow_taper" -p no:cacheprovider > gpurun_out/r2b_taper.log 2>&1; echo "taper rc=$?"
grep -n "AssertionError" gpurun_out/r2b_taper.log | cut -c1-1500
python __graft_entry__.py --smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2b_smoke.log
