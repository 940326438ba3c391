#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/parity
export GLB_SE_3PASS=1
timeout 300 python tests/calibrate_tf32_bounds.py > gpurun_out/r2_reference_on_b200_deviation.json 2> gpurun_out/r2d_calib.err; echo "calib rc=$?"; tail -3 gpurun_out/r2d_calib.err
GLB_DUMP_PARITY=gpurun_out/parity timeout 600 python -m pytest tests/test_cfg2_fullwidth.py -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2d_cfg2.log 2>&1; echo "cfg2 rc=$?"
grep -n "^E  \|passed\|failed" gpurun_out/r2d_cfg2.log | cut -c1-300 | head -12
timeout 600 python -m pytest tests/test_zz_gpu_widen.py -m gpu -q -k "grow" -p no:cacheprovider > gpurun_out/r2d_grow.log 2>&1; echo "grow rc=$?"; tail -3 gpurun_out/r2d_grow.log
timeout 200 python tools/check_tc.py fprop dgrad > gpurun_out/r2d_tc_base.txt 2>&1; echo "base rc=$?"
GLB_FPROP_HALO=3 GLB_FPROP_ONEBOX=1 timeout 200 python tools/check_tc.py fprop > gpurun_out/r2d_tc_onebox.txt 2>&1; echo "onebox rc=$?"
GLB_FPROP_HALO=3 timeout 200 python tools/check_tc.py fprop > gpurun_out/r2d_tc_halo1cta.txt 2>&1; echo "halo1cta rc=$?"
GLB_FPROP_H1=1 timeout 200 python tools/check_tc.py fprop dgrad > gpurun_out/r2d_tc_h1.txt 2>&1; echo "h1 rc=$?"
GLB_FPROP_H1=2 timeout 200 python tools/check_tc.py fprop > gpurun_out/r2d_tc_h1_256.txt 2>&1; echo "h1_256 rc=$?"
for f in base onebox halo1cta h1 h1_256; do echo "== $f"; grep -E "64x64|128x128" gpurun_out/r2d_tc_$f.txt | cut -c1-110; done
unset GLB_SE_3PASS
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "style or instance" -p no:cacheprovider > gpurun_out/r2d_style.log 2>&1; echo "style (cluster kernels) rc=$?"; tail -8 gpurun_out/r2d_style.log | cut -c1-300
timeout 300 python tools/glue_bw.py > gpurun_out/r2d_glue_cluster.txt 2>&1; echo "glue rc=$?"; grep -i "style" gpurun_out/r2d_glue_cluster.txt | head
GLB_SE_3PASS=1 timeout 300 python tools/glue_bw.py > gpurun_out/r2d_glue_3pass.txt 2>&1; grep -i "style" gpurun_out/r2d_glue_3pass.txt | head
