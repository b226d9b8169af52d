#!/bin/bash
# GPU batch 8 (2 GPUs): next-frame raycast prefetch on top of the frame overlap — tests, N = 1 at 4K, 2-GPU proxy at the 8-GPU slab size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "overlap or reprojection or frame_hash" > gpurun_out/pytest_b8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b8.log
B="--steps 16 --warmup 4 --no-cpu-baseline --no-fast-line"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py $B --no-ref-gpu > gpurun_out/bench_r2h_n1.json 2> gpurun_out/bench_r2h_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2h_n1.json")); print("n1 %.1f Mpix/s %.3f ms e2e %.1f hash %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"]), {k:round(x["ms_per_launch"],3) for k,x in d["kernels"].items()})
PY
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 $B --no-frame-hash "$@" > gpurun_out/bench_r2h_$tag.json 2> gpurun_out/bench_r2h_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2h_$tag.json")); print("%-16s %8.1f Mpix/s %7.3f ms e2e %8.1f timed_out %s"%("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], d["halo_wait_timed_out"]))
except Exception as e: print("$tag failed", e)
PY
}
run h544_ov1 --height 544 --overlap 1
run h544_ov0 --height 544 --overlap 0
run h2160_ov1 --overlap 1
