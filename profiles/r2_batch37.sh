#!/bin/bash
# GPU batch 37: RIS loop with deferred selection (default) against the copying form (copy), at 4 blocks per SM (ct4), with three
# light records in flight (rb3), with the shared-reciprocal vector division (d3); div3_shared against `/` bit for bit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in nofmad fmad; do timeout 300 profiles/microbench/div3_check_$m 0x200000000 > gpurun_out/div3_check_$m.txt 2>&1; echo "div3_check_$m rc=$? $(cat gpurun_out/div3_check_$m.txt)"; done
for v in "" copy ct4 rb3 d3 ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b37_$v.json 2> gpurun_out/bench_b37_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b37_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
