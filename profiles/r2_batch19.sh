#!/bin/bash
# GPU batch 19: whole GPU suite on the own-triangle build + pipelined read-back through two device images; bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_b19.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_b19.log
timeout 900 python bench.py > gpurun_out/bench_b19_n1.json 2> gpurun_out/bench_b19_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_b19_n1.json")); print("n1 %.1f Mpix/s %.3f ms e2e %.1f hash %s ref_gpu %s fast %s issue %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"],d.get("ref_gpu",{}).get("value"),(d.get("fast_math") or {}).get("value"), d["roofline_issue"]["frame_frac"]), {k:(round(x["ms_per_launch"],3), x.get("issue_frac")) for k,x in d["kernels"].items()})
PY
timeout 600 python bench.py --readback sync --no-cpu-baseline --no-ref-gpu --no-fast-line > gpurun_out/bench_b19_n1_sync.json 2> gpurun_out/bench_b19_n1_sync.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b19_n1_sync.json')); print('sync readback: value', d['value'], 'e2e', d['e2e']['value'])"
