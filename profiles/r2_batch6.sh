#!/bin/bash
# GPU batch 6: full GPU test suite on the current build + configs 2-4 after the per-lane walk-step loops + ncu of their kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_b6.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_b6.log
for c in 06 08 09; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/bench_r2f_cfg$c.json 2> gpurun_out/bench_r2f_cfg$c.err; echo "cfg$c rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_r2f_cfg$c.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'],d['e2e']['value'],d['cpu_baseline']['value'])")"; done
for c in 06 08 09; do
  K='regex:k_(ao|path_trace)'
  ncu --set full --clock-control none --import-source on -k "$K" -s 4 -c 1 -f -o gpurun_out/prof_r2f_cfg$c python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2f_cfg$c.log 2>&1
  echo "ncu cfg$c rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -4
