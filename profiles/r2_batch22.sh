#!/bin/bash
# GPU batch 22: whole GPU suite with the fused-frame reprojection and the edge-case tests; bench line (no regression of the in-place kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_b22.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_b22.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b22_n1.json 2> gpurun_out/bench_b22_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_b22_n1.json")); print("n1 %.1f Mpix/s %.3f ms e2e %.1f hash %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
