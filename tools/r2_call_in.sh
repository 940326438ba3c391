#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_input_pipeline.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300
echo "== grouped kernel"; timeout 300 python tools/input_bw.py 2>&1 | tee gpurun_out/r2_input_pipeline_bw.txt | tail -8
echo "== v1 kernel"; GLB_INPUT_V1=1 timeout 300 python tools/input_bw.py 2>&1 | tail -8
