#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/step_calls.py --filter bn_act_bwd_apply > gpurun_out/g_calls_device.txt 2>&1; cat gpurun_out/g_calls_device.txt
timeout 300 python tools/step_calls.py --conn host --filter bn_act_bwd_apply > gpurun_out/g_calls_host.txt 2>&1; cat gpurun_out/g_calls_host.txt
timeout 300 python tools/step_calls.py --filter "" > gpurun_out/g_calls_all.txt 2>&1; tail -1 gpurun_out/g_calls_all.txt
timeout 300 python tools/ffn_trace.py --dgrad > gpurun_out/g_dgrad_trace.json 2> gpurun_out/g_dgrad_trace.err; echo "trace rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/g_dgrad_trace.json"))
    print(json.dumps(d.get("steady_state"), indent=1))
    for r in d["mma"][24:40]: print(r)
    for r in d["transform_g0"][4:8]: print("g0", r)
    for r in d["transform_g1"][4:8]: print("g1", r)
except Exception as e:
    print("trace parse failed", e)
PY
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/g_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/g_suite.log
