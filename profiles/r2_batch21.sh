#!/bin/bash
# GPU batch 21: edge-case tests; final lines of configs 2-4 and of the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_edges.py -q -m gpu > gpurun_out/pytest_b21.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_b21.log
for c in 06 08 09; do timeout 600 python bench.py --config $c > gpurun_out/bench_final_cfg$c.json 2> gpurun_out/bench_final_cfg$c.err; echo "cfg$c rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_final_cfg$c.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'],d['cpu_baseline']['value'],d['roofline']['kernel'],d['roofline']['frac'])")"; done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "ref arm rc=$?"; cat gpurun_out/bench_final_reference.json | head -c 600
