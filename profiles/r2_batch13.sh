#!/bin/bash
# GPU batch 13: pooled persistent closest-hit kernel for the path tracers' camera and bounce rays
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -k "path_tracers or config3 or config4 or launch_by_name" > gpurun_out/pytest_b13.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_b13.log
for c in 08 09; do for pc in 2 1 0; do CRT_POOLED_CLOSEST=$pc timeout 400 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2m_cfg${c}_pc$pc.json 2> gpurun_out/bench_r2m_cfg${c}_pc$pc.err; echo "cfg$c pooled_closest=$pc rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_r2m_cfg${c}_pc$pc.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'],{k:v['ms_per_frame'] for k,v in d['kernels'].items()})")"; done; done
