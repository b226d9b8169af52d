#!/bin/bash
# GPU batch 2 of round 2: far-first shadow order + register-resident walk state; FFMA2 variant A/B; ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bvh_equals or raycast or refit or fused_frame_bit_exact or full_size_band or path_tracers or ao06" > gpurun_out/pytest_b2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b2.log
B="--steps 8 --warmup 3 --no-ref-gpu --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err; echo "bench rc=$?"
CRT_LIB_VARIANT=ffma2 timeout 600 python bench.py $B > gpurun_out/bench_r2b_ffma2.json 2> gpurun_out/bench_r2b_ffma2.err; echo "bench ffma2 rc=$?"
python - <<'PY'
import json
for t in ("n1","ffma2"):
    try:
        d=json.load(open("gpurun_out/bench_r2b_%s.json"%t))
        print(t, d["value"], d["ms_per_step"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()}, d["frame_hash"]["value"])
    except Exception as e: print(t, "failed", e)
PY
for c in 06 08 09; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_cfg$c.json 2> gpurun_out/bench_r2b_cfg$c.err; echo "cfg$c rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_r2b_cfg$c.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'])")"; done
bash profiles/capture.sh r2b > gpurun_out/capture_r2b.log 2>&1; echo "capture rc=$?"; tail -3 gpurun_out/capture_r2b.log
