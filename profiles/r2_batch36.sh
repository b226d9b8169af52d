#!/bin/bash
# GPU batch 36 (2 GPUs): multi-GPU slab tests and the 2-GPU bench line on the build with seeded primary rays
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "slabs_on_gpus" > gpurun_out/pytest_b36.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b36.log
run() { # n, tag, extra args
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 16 --warmup 4 --no-cpu-baseline --no-fast-line "$@" > gpurun_out/bench_b36_$tag.json 2> gpurun_out/bench_b36_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b36_$tag.json"))
    print("%-12s %8.1f Mpix/s %7.3f ms e2e %8.1f hash %s timed_out %s slab_ms %s"%("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("frame_hash") or {}).get("value"), d["halo_wait_timed_out"], d["slab_kernel_ms"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
except Exception as e: print("$tag failed", e)
PY
}
run 2 n2
# the 8-GPU slab size on two GPUs (3840 x 544: 272 rows each), as in batch 4
run 2 n2_h544 --height 544 --no-frame-hash
