#!/bin/bash
# GPU batch 39 (2 GPUs): temporal reprojection across row slabs (exchange mode: every rank gathers every slab's history rows),
# launch-list and fused form, against the one-slab frames with the same moving camera
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x -k "slabs_on_gpus and reproject" > gpurun_out/pytest_b39.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_b39.log
