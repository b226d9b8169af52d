#!/bin/bash
# GPU batch 15: own-triangle pre-test at ray emission (visibility-reuse rays of the fused frame, shadow rays of 08_nee / 09_ris)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_b15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_b15.log
for v in "" noown; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_b15_n1_$v.json 2> gpurun_out/bench_b15_n1_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b15_n1_$v.json")); print("n1[$v] %.1f Mpix/s %.3f ms e2e %.1f hash %s rays %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"],d["rays"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
  for c in 08 09; do CRT_LIB_VARIANT=$v timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b15_cfg${c}_$v.json 2> gpurun_out/bench_b15_cfg${c}_$v.err; echo "cfg$c[$v] rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_b15_cfg${c}_$v.json'));print(d['value'],d['ms_per_step'],d['grays_per_s'],d['rays'],{k:v['ms_per_frame'] for k,v in d['kernels'].items()})")"; done
done
