#!/usr/bin/env bash
timeout 150 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 | cut -c1-200
