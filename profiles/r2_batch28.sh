#!/bin/bash
# GPU batch 28: light pdf folded into the light table, depths of the spatial pass computed once, sincosf; whole GPU suite on the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
profiles/microbench/sincos_check_fmad; profiles/microbench/sincos_check_nofmad
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_b28.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_b28.log
for v in "" sincos ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b28_$v.json 2> gpurun_out/bench_b28_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b28_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
CRT_LIB_VARIANT=sincos timeout 900 python -m pytest tests -q -m gpu -x -k "fused or tolerance or default_math" > gpurun_out/pytest_b28_sincos.log 2>&1; echo "pytest[sincos] rc=$?"; tail -3 gpurun_out/pytest_b28_sincos.log
