#!/bin/bash
# GPU batch 32: primary rays seeded with the pixel's previous answer (CRT_RAYCAST_HINT, A/B by environment in one build);
# fewer resident tracer blocks per SM beside the overlapping frame (CRT_RESOLVE_BLOCKS_PER_SM / CRT_TRACE_BLOCKS_PER_SM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "raycast or frame_hash or overlap or fused" > gpurun_out/pytest_b32.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b32.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b32_$tag.json 2> gpurun_out/bench_b32_$tag.err; echo "bench[$tag] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b32_$tag.json")); print("[$tag]: %.1f Mpix/s %.3f ms e2e %.1f hash %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
}
run hint1 CRT_RAYCAST_HINT=1
run hint0 CRT_RAYCAST_HINT=0
run hint1b CRT_RAYCAST_HINT=1
run hint0b CRT_RAYCAST_HINT=0
run rb4 CRT_RESOLVE_BLOCKS_PER_SM=4
run rb6 CRT_RESOLVE_BLOCKS_PER_SM=6
run tb6 CRT_TRACE_BLOCKS_PER_SM=6
run rb6tb6 CRT_RESOLVE_BLOCKS_PER_SM=6 CRT_TRACE_BLOCKS_PER_SM=6
