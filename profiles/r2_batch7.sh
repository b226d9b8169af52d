#!/bin/bash
# GPU batch 7: per-lane ray loops (entry-mask ballots, rays started 8 at a time) against the lockstep loops; slab probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "reprojection or config2 or config4 or launch_by_name" > gpurun_out/pytest_b7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b7.log
for v in default lockstep; do
  if [ $v = default ]; then unset CRT_LIB_VARIANT; else export CRT_LIB_VARIANT=$v; fi
  for c in 06 08 09; do timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2g_${v}_cfg$c.json 2> gpurun_out/bench_r2g_${v}_cfg$c.err; echo "$v cfg$c rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_r2g_${v}_cfg$c.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'])")"; done
done
unset CRT_LIB_VARIANT
timeout 900 python profiles/slab_probe.py 2 4 8 > gpurun_out/slab_probe.json 2> gpurun_out/slab_probe.err; echo "probe rc=$?"; cat gpurun_out/slab_probe.json
