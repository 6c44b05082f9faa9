#!/usr/bin/env bash
# GPU session D: FFN forward phase trace; new tests (corrcoef, d128 golden fused tail); smoke
set -u
mkdir -p gpurun_out
timeout 300 python tools/ffn_trace.py > gpurun_out/d_ffn_trace.json 2> gpurun_out/d_ffn_trace.err; echo "trace rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/d_ffn_trace.json"))
    print(json.dumps(d.get("steady_state"), indent=1))
    for r in d["mma"][16:32]: print(r)
    for r in d["transform_g0"][4:10]: print("g0", r)
    for r in d["transform_g1"][4:10]: print("g1", r)
    for r in d["producer"][16:28]: print("P", r)
except Exception as e:
    print("trace parse failed", e)
PY
timeout 600 python -m pytest tests -x -q -m gpu -k "corrcoef or d128 or ffn" > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/d_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/d_smoke.log
