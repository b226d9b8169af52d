#!/bin/bash
# GPU batch 30: one queue-append atomic per block instead of per warp (k_resolve_fast, k_candidate_temporal)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" bpr bpc bpb ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b30_$v.json 2> gpurun_out/bench_b30_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b30_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
