#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + full ncu capture of one steady-state frame of bench.py.
# usage: profiles/capture.sh <tag> [kernel-regex] [launches-per-frame]      outputs under gpurun_out/
TAG=${1:-r1}
K=${2:-'regex:k_(raycast|generate_candidate|temporal|save_temporal|spatial|resolve|tone_mapping|trace_shadow_queue|candidate_temporal|spatial_fast|resolve_fast)'}
N=${3:-9}
MODE=${4:-fused}
# every launch of frames 2-3 with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s $N -c $((2*N)) --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-frame-hash --no-fast-line --mode $MODE > gpurun_out/ncu_launches_${TAG}.log 2>&1
# one full capture of frame 2's kernels
ncu --set full --clock-control none --import-source on -k "$K" -s $N -c $N -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-frame-hash --no-fast-line --mode $MODE > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -5
