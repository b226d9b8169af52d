#!/bin/bash
# GPU batch 10: persistent tracer tuning after this round's changes (resident blocks, refill threshold), the fixed test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "resolve_reuse" > gpurun_out/pytest_b10.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_b10.log
B="--steps 8 --warmup 3 --no-ref-gpu --no-cpu-baseline --no-fast-line --no-frame-hash"
for v in default mb6 mb10 rf16 rf20 rf28; do
  if [ $v = default ]; then unset CRT_LIB_VARIANT; else export CRT_LIB_VARIANT=$v; fi
  timeout 300 python bench.py $B > gpurun_out/bench_r2j_$v.json 2> gpurun_out/bench_r2j_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2j_$v.json"))
    print("%-8s 4K   %7.1f Mpix/s %6.3f ms "%("$v", d["value"], d["ms_per_step"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items() if "trace" in k})
except Exception as e: print("$v failed", e)
PY
  timeout 300 python bench.py $B --height 272 > gpurun_out/bench_r2j_${v}_h272.json 2> gpurun_out/bench_r2j_${v}_h272.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2j_${v}_h272.json"))
    print("%-8s h272 %7.1f Mpix/s %6.3f ms "%("$v", d["value"], d["ms_per_step"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items() if "trace" in k})
except Exception as e: print("$v h272 failed", e)
PY
done
