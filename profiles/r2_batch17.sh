#!/bin/bash
# GPU batch 17: prefetch of the next node / the first triangle record ahead of the pooled triangle phase (persistent tracer)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" pfn pft pfb ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b17_$v.json 2> gpurun_out/bench_b17_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b17_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
