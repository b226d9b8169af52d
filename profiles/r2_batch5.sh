#!/bin/bash
# GPU batch 5 (8 GPUs): the scaling points the driver measures, with and without frame overlap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # n, tag, extra args
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 16 --warmup 4 --no-cpu-baseline --no-fast-line "$@" > gpurun_out/bench_r2e_$tag.json 2> gpurun_out/bench_r2e_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2e_$tag.json"))
    print("%-12s %8.1f Mpix/s %7.3f ms e2e %8.1f hash %s timed_out %s slab_ms %s"%("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("frame_hash") or {}).get("value"), d["halo_wait_timed_out"], d["slab_kernel_ms"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
except Exception as e: print("$tag failed", e)
PY
}
run 8 n8_ov1 --overlap 1
run 8 n8_ov0 --overlap 0 --no-frame-hash
run 4 n4_ov1 --overlap 1
