#!/bin/bash
# GPU batch 3 of round 2: A/B of the walk variants (per-thread walk is sensitive to code generation), frame overlap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "overlap or frame_hash" > gpurun_out/pytest_b3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b3.log
B="--steps 8 --warmup 3 --no-ref-gpu --no-cpu-baseline --no-fast-line --no-frame-hash"
for v in default oneobj oneobj_ffma2 ballot ballot_ffma2 ffma2; do
  for ov in 0 1; do
    if [ $v = default ]; then unset CRT_LIB_VARIANT; else export CRT_LIB_VARIANT=$v; fi
    timeout 300 python bench.py $B --overlap $ov > gpurun_out/bench_r2c_${v}_ov$ov.json 2> gpurun_out/bench_r2c_${v}_ov$ov.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2c_${v}_ov$ov.json"))
    print("%-14s ov$ov %7.1f Mpix/s %6.3f ms  e2e %7.1f "%("$v", d["value"], d["ms_per_step"], d["e2e"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
except Exception as e: print("$v ov$ov failed", e)
PY
  done
  for c in 06 08 09; do timeout 300 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c_${v}_cfg$c.json 2> gpurun_out/bench_r2c_${v}_cfg$c.err; echo "   $v cfg$c $(python -c "import json;d=json.load(open('gpurun_out/bench_r2c_${v}_cfg$c.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'])")"; done
done
