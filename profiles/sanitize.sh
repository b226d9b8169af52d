#!/bin/bash
# Runs on the GPU box (under gpurun): compute-sanitizer over smoke() (launch list + fused frame, 160x90, bit-exact vs the
# oracle) and over the one-GPU slab-group frame (three slabs, three streams, signal/wait flags, halo rows stored through
# plain pointers by the producing kernels: the protocol of csrc/slab_p2p.cu).  Logs land in gpurun_out/sanitize_*.log;
# the summaries are committed under profiles/r2/.
SAN=/usr/local/cuda/bin/compute-sanitizer
SMOKE='import __graft_entry__ as g; g.smoke()'
for tool in memcheck initcheck racecheck synccheck; do
  timeout 900 $SAN --tool $tool --print-limit 20 --error-exitcode 0 python -c "$SMOKE" > gpurun_out/sanitize_smoke_$tool.log 2>&1
  echo "smoke/$tool rc=$? : $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/sanitize_smoke_$tool.log | tr '\n' ' ')"
done
timeout 1500 $SAN --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "slab_group_on_one_gpu" > gpurun_out/sanitize_slabgroup_memcheck.log 2>&1
echo "slabgroup/memcheck rc=$? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_slabgroup_memcheck.log | tr '\n' ' ')"
