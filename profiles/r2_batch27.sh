#!/bin/bash
# GPU batch 27: Triangle AoS gathers of the per-pixel kernels past the L1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" trinoalloc ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b27_$v.json 2> gpurun_out/bench_b27_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b27_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
