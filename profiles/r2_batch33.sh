#!/bin/bash
# GPU batch 33: triangle-postponing threshold of the per-thread walk with seeded primary rays (CRT_POSTPONE; default 0.25)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b33_$tag.json 2> gpurun_out/bench_b33_$tag.err; echo "bench[$tag] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b33_$tag.json")); print("[$tag]: %.1f Mpix/s %.3f ms e2e %.1f hash %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
}
run p000 CRT_POSTPONE=0
run p015 CRT_POSTPONE=0.15
run p025 CRT_POSTPONE=0.25
run p040 CRT_POSTPONE=0.4
run p060 CRT_POSTPONE=0.6
