#!/usr/bin/env bash
# full validation as the driver runs it: GPU suite, smoke, default bench (both arms)
set -u
mkdir -p gpurun_out/parity
GLB_DUMP_PARITY=gpurun_out/parity timeout 1200 python -m pytest tests -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log | cut -c1-300
cat gpurun_out/parity/resnet_nets_*.json
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2l_smoke.log | cut -c1-300
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench.err | cut -c1-300
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2l_ref.json 2> gpurun_out/r2l_ref.err ) 2>&1 | grep real; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2l_bench.json").read().splitlines() if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","dtype","config")})
print("e2e", d["e2e"], "launches", d["gpu_launches"], "clocks", d["clocks"])
r=d["roofline"]; print("roofline", {k:r[k] for k in ("achieved","frac","frac_of_tf32_mma_peak","traffic")}, r.get("split_by_roofline"))
g=d["roofline_glue"]; print("glue", g["achieved"], g["frac"])
print("bf16", d.get("opt_in_bf16_operands")); print("eager", d.get("torch_eager_b200")); print("cpu", d.get("cpu_baseline"))
r=json.loads([l for l in open("gpurun_out/r2l_ref.json").read().splitlines() if l.startswith("{")][-1])
print("reference arm", {k:r.get(k) for k in ("value","ms_per_step","config","steps")})
PY
