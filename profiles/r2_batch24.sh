#!/bin/bash
# GPU batch 24: cache policies — 256-bit gathers past the L1 or with a 128-byte L2 fetch; triangle records streamed, nodes kept
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" noalloc l2128 trina trief nodeel both ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b24_$v.json 2> gpurun_out/bench_b24_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b24_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
