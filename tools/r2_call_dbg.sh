#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "upconv_family" -p no:cacheprovider 2>&1 | tail -30 | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep FAILED | cut -c1-200
