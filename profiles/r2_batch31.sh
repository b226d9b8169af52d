#!/bin/bash
# GPU batch 31: the tracer's refill with one atomic per chunk of rays instead of one per refill; k_resolve_fast with one append atomic per block (default now)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" fc16 fc32 fc64 ""; do
  CRT_LIB_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b31_$v.json 2> gpurun_out/bench_b31_$v.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b31_$v.json")); print("[$v]: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
for v in "" fc16 fc32; do
  CRT_LIB_VARIANT=$v timeout 600 python profiles/slab_probe.py 8 > gpurun_out/slab_probe_b31_$v.json 2> gpurun_out/slab_probe_b31_$v.err; echo "probe[$v] rc=$?"
  python - <<PY
import json
for line in open("gpurun_out/slab_probe_b31_$v.json"):
    d=json.loads(line); print("[$v] slabs", d["slabs"], d.get("sum_ms"), d.get("kernels_ms") or d.get("kernels_ms_summed_over_slabs"), d.get("ratio_sum"))
PY
done
