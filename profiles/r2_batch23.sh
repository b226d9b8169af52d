#!/bin/bash
# GPU batch 23 (2 GPUs): slim halo mirroring (plane 0 only for records a neighbour cannot select) at the 8-GPU slab size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "slabs_on_gpus or slab_group" > gpurun_out/pytest_b23.log 2>&1; echo "pytest slabs rc=$?"; tail -3 gpurun_out/pytest_b23.log
run() { # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 16 --warmup 4 --no-cpu-baseline --no-fast-line --no-frame-hash "$@" > gpurun_out/bench_b23_$tag.json 2> gpurun_out/bench_b23_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b23_$tag.json"))
    print("%-22s %8.1f Mpix/s %7.3f ms e2e %8.1f  slab_ms %s timed_out %s"%("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], d["slab_kernel_ms"], d["halo_wait_timed_out"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
except Exception as e: print("$tag failed", e)
PY
}
for slim in 0 1 0 1; do CRT_HALO_SLIM=$slim run h544_slim$slim --height 544; done
CRT_HALO_SLIM=1 run h2160_slim1
