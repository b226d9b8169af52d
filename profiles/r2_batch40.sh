#!/bin/bash
# GPU batch 40: the GPU suite, smoke and the default bench line on the final tree of the round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final.log)"
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final_n1.json")); print("n1 %.1f Mpix/s %.3f ms e2e %.1f hash %s ref_gpu %s fast %s issue %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["frame_hash"]["value"],d.get("ref_gpu",{}).get("value"),(d.get("fast_math") or {}).get("value"), d["roofline_issue"]["frame_frac"]), {k:(round(x["ms_per_launch"],3), x.get("issue_frac")) for k,x in d["kernels"].items()})
PY
