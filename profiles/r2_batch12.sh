#!/bin/bash
# GPU batch 12: examples 07-09 as a wavefront against the single kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -k "path_tracers or config3 or config4 or launch_by_name or ao06 or config2" > gpurun_out/pytest_b12.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_b12.log
for c in 08 09; do for wf in 1 0; do CRT_WAVEFRONT=$wf timeout 400 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2l_cfg${c}_wf$wf.json 2> gpurun_out/bench_r2l_cfg${c}_wf$wf.err; echo "cfg$c wf=$wf rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_r2l_cfg${c}_wf$wf.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'],{k:v['ms_per_frame'] for k,v in d['kernels'].items()})")"; done; done
