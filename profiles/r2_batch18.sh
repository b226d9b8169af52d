#!/bin/bash
# GPU batch 18: pooled triangle phase with carry-over (only well-filled rounds are run; leftover pairs wait for the next iteration)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CRT_LIB_VARIANT=tc16 timeout 900 python -m pytest tests -q -m gpu -x -k "fused or full_size or hash or reference_launch" > gpurun_out/pytest_b18.log 2>&1; echo "pytest[tc16] rc=$?"; tail -3 gpurun_out/pytest_b18.log
for v in "" tc33 tc24 tc16 tc8 ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b18_$v.json 2> gpurun_out/bench_b18_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b18_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
