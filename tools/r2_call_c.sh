#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/parity
python tests/calibrate_tf32_bounds.py > gpurun_out/r2_reference_on_b200_deviation.json 2> gpurun_out/r2c_calib.err; echo "calib rc=$?"; tail -3 gpurun_out/r2c_calib.err
GLB_DUMP_PARITY=gpurun_out/parity python -m pytest tests/test_cfg2_fullwidth.py -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2c_cfg2.log 2>&1; echo "cfg2 rc=$?"
grep -n "^E  \|passed\|failed" gpurun_out/r2c_cfg2.log | cut -c1-400 | head -20
python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow" -p no:cacheprovider > gpurun_out/r2c_taper.log 2>&1; echo "grow rc=$?"; tail -3 gpurun_out/r2c_taper.log
