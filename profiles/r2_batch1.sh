#!/bin/bash
# GPU batch 1 of round 2: new parity tests, the bench line with the comparator, configs 2-4, HIPRT comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_configs.py -q -m gpu -s -x > gpurun_out/pytest_configs.log 2>&1; echo "pytest configs rc=$?"; tail -5 gpurun_out/pytest_configs.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "fast_math_mode or smoke or launch_by_name or fused_frame_bit_exact" > gpurun_out/pytest_math.log 2>&1; echo "pytest math rc=$?"; grep -E "mean relative|passed|failed" gpurun_out/pytest_math.log | tail -5
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2a_n1.json 2> gpurun_out/bench_r2a_n1.err; echo "bench rc=$?"; head -c 1500 gpurun_out/bench_r2a_n1.json; echo
for c in 06 08 09; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/bench_r2a_cfg$c.json 2> gpurun_out/bench_r2a_cfg$c.err; echo "bench cfg$c rc=$?"; head -c 600 gpurun_out/bench_r2a_cfg$c.json; echo; done
timeout 900 python oracle/ref_gpu/run_ref_gpu.py --frames 61 --warmup 3 --compare --compare-radiance > gpurun_out/ref_hiprt_r2.json 2> gpurun_out/ref_hiprt_r2.log; echo "ref rc=$?"; cat gpurun_out/ref_hiprt_r2.json
