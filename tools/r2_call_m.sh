#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log | cut -c1-300
for v in "A=1" "GLB_NO_ARENA=1"; do env $v timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2m_bench_$v.json 2> gpurun_out/r2m_bench_$v.err; echo "bench $v rc=$?"; tail -2 gpurun_out/r2m_bench_$v.err | cut -c1-200; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_bench_*.json")):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1]); print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median")}, d["e2e"]["value"])
PY
