#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "corrcoef or d128 or ffn or tail_fused" > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/e_tests.log
timeout 300 python tools/ffn_bench.py --only-fused > gpurun_out/e_ffn_bench.json 2> gpurun_out/e_ffn_bench.err; cat gpurun_out/e_ffn_bench.json
timeout 300 python tools/ffn_trace.py > gpurun_out/e_ffn_trace.json 2> gpurun_out/e_ffn_trace.err; echo "trace rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/e_ffn_trace.json"))
    print(json.dumps(d.get("steady_state"), indent=1))
    for r in d["mma"][16:28]: print(r)
    for r in d["transform_g0"][4:8]: print("g0", r)
    for r in d["transform_g1"][4:8]: print("g1", r)
except Exception as e:
    print("trace parse failed", e)
PY
