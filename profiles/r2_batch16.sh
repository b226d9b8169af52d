#!/bin/bash
# GPU batch 16: two length classes in the ray queue (long rays fetched first): one GPU at 4K and the slab probe at the 8-GPU slab size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -k "fused or slab or overlap or full_size or hash or config4" > gpurun_out/pytest_b16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b16.log
for s in 0 16 8 32; do
  CRT_QUEUE_SPLIT=$s timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b16_split$s.json 2> gpurun_out/bench_b16_split$s.err; echo "bench[split $s] rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b16_split$s.json")); print("split $s: %.1f Mpix/s %.3f ms hash %s"%(d["value"],d["ms_per_step"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
done
for s in 0 16 8; do
  CRT_QUEUE_SPLIT=$s timeout 600 python profiles/slab_probe.py 8 > gpurun_out/slab_probe_b16_split$s.json 2> gpurun_out/slab_probe_b16_split$s.err; echo "probe[split $s] rc=$?"
  python - <<PY
import json
for line in open("gpurun_out/slab_probe_b16_split$s.json"):
    d=json.loads(line); print("split $s slabs", d["slabs"], d.get("sum_ms"), d.get("kernels_ms") or d.get("kernels_ms_summed_over_slabs"), d.get("ratio_sum"))
PY
done
