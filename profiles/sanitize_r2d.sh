#!/bin/bash
# Runs on the GPU box (under gpurun): compute-sanitizer over the code that changed in the second half of round 2 — the
# own-triangle pre-test, the hole records of the wavefront 09_ris, the fused frame with reprojection — through smoke() and
# the small edge-case / reprojection tests.  Logs: gpurun_out/sanitize_r2d_*.log; summary committed under profiles/r2/sanitizer/.
SAN=/usr/local/cuda/bin/compute-sanitizer
SMOKE='import __graft_entry__ as g; g.smoke()'
for tool in memcheck initcheck racecheck; do
  timeout 900 $SAN --tool $tool --print-limit 20 --error-exitcode 0 python -c "$SMOKE" > gpurun_out/sanitize_r2d_smoke_$tool.log 2>&1
  echo "smoke/$tool rc=$? : $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/sanitize_r2d_smoke_$tool.log | tr '\n' ' ')"
done
for tool in memcheck racecheck; do
  timeout 1500 $SAN --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_edges.py tests/test_gpu_configs.py -x -q -m gpu \
    -k "ragged_image_sizes_examples or single_light or (ragged_image_sizes_fused and size1) or temporal_reprojection" > gpurun_out/sanitize_r2d_edges_$tool.log 2>&1
  echo "edges/$tool rc=$? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_r2d_edges_$tool.log | tr '\n' ' ')"
done
