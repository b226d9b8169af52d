#!/bin/bash
# GPU batch 4 (2 GPUs): proxy for the 8-GPU slab size — a 3840x544 frame on two GPUs (272 rows each, what a GPU renders of the
# 4K frame at N = 8): frame overlap x slabs per GPU; then the same at 4K on two GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 16 --warmup 4 --no-cpu-baseline --no-fast-line --no-frame-hash "$@" > gpurun_out/bench_r2d_$tag.json 2> gpurun_out/bench_r2d_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2d_$tag.json"))
    print("%-22s %8.1f Mpix/s %7.3f ms e2e %8.1f  slab_ms %s timed_out %s"%("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], d["slab_kernel_ms"], d["halo_wait_timed_out"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
except Exception as e: print("$tag failed", e)
PY
}
for ov in 0 1; do for sub in 1 2; do run h544_ov${ov}_sub${sub} --height 544 --overlap $ov --sub $sub; done; done
for ov in 0 1; do run h2160_ov${ov}_sub1 --overlap $ov --sub 1; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "slabs_on_gpus" > gpurun_out/pytest_b4.log 2>&1; echo "pytest slabs rc=$?"; tail -3 gpurun_out/pytest_b4.log
