#!/usr/bin/env bash
# Round-2 first GPU call: un-xfailed suite with tracebacks, measured MMA peaks, the reference through stock PyTorch on the B200,
# and the pending A/B switches.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxXfE -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
./tools/micro/mma_peak > gpurun_out/r2a_mma_peak.txt 2>&1; cat gpurun_out/r2a_mma_peak.txt
python bench.py --impl reference --ref-dev cuda --steps 10 --warmup 3 > gpurun_out/r2a_ref_cuda.json 2> gpurun_out/r2a_ref_cuda.err; echo "ref cuda rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
GLB_BATCH_D=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_batchd.json 2> gpurun_out/r2a_bench_batchd.err; echo "bench batch_d rc=$?"
GLB_RESNET_GRAPHS=1 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg5_graphs.json 2> gpurun_out/r2a_bench_cfg5_graphs.err; echo "cfg5 graphs rc=$?"
for f in r2a_ref_cuda r2a_bench r2a_bench_batchd r2a_bench_cfg5_graphs; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
