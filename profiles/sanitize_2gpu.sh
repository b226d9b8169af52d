#!/bin/bash
# Runs on a 2-GPU box (gpurun --gpus 2): compute-sanitizer memcheck over the two-process row-slab frame — halo rows stored by the
# producing kernels into the neighbour's buffers through cudaIpc peer pointers over NVLink, signal/wait flags across processes
# (csrc/slab_p2p.cu) — tests/gpu_slab_worker.py, mode "fused" (direct peer stores), 3 frames, bit-identity check included.
SAN=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
SLAB_MODE=fused timeout 1500 $SAN --tool memcheck --target-processes all --print-limit 20 --error-exitcode 0 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tests/gpu_slab_worker.py \
  > gpurun_out/sanitize_2gpu_memcheck.log 2>&1
echo "2gpu/memcheck rc=$? : $(grep -E 'ERROR SUMMARY|GPU_SLABS' gpurun_out/sanitize_2gpu_memcheck.log | sort | uniq -c | tr '\n' ' ')"
tail -5 gpurun_out/sanitize_2gpu_memcheck.log
