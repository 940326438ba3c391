#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12 | cut -c1-400
timeout 300 python tools/check_upconv.py > gpurun_out/r2s_check_upconv.txt 2>&1; cat gpurun_out/r2s_check_upconv.txt | tail -20
