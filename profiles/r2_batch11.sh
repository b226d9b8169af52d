#!/bin/bash
# GPU batch 11: 06_ao as a wavefront (emit / persistent tracer / finish) against the single kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "ao06 or config2 or launch_by_name_round or smoke" > gpurun_out/pytest_b11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b11.log
for wf in 1 0; do CRT_WAVEFRONT=$wf timeout 300 python bench.py --config 06 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2k_cfg06_wf$wf.json 2> gpurun_out/bench_r2k_cfg06_wf$wf.err; echo "wf=$wf rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_r2k_cfg06_wf$wf.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'],d['kernels'])")"; done
