#!/usr/bin/env bash
# First GPU call of a round: everything that was written without hardware at hand.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
# Outputs go to gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxX -p no:cacheprovider > gpurun_out/fc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/fc_pytest.log
tail -40 gpurun_out/fc_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/fc_bench.json 2> gpurun_out/fc_bench.err; echo "bench rc=$?"
GLB_BATCH_D=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/fc_bench_batchd.json 2> gpurun_out/fc_bench_batchd.err; echo "bench batch_d rc=$?"
python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fc_bench_cfg5.json 2> gpurun_out/fc_bench_cfg5.err; echo "cfg5 rc=$?"
GLB_RESNET_GRAPHS=1 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fc_bench_cfg5_graphs.json 2> gpurun_out/fc_bench_cfg5_graphs.err; echo "cfg5 graphs rc=$?"
python bench.py --impl reference --ref-dev cuda --steps 10 --warmup 3 > gpurun_out/fc_ref_cuda.json 2> gpurun_out/fc_ref_cuda.err; echo "ref cuda rc=$?"
python tools/input_bw.py > gpurun_out/fc_input_bw.txt 2>&1
for f in fc_bench fc_bench_batchd fc_bench_cfg5 fc_bench_cfg5_graphs fc_ref_cuda; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
cat gpurun_out/fc_input_bw.txt
