#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/l_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/l_suite.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/l_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/l_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "loss", d["config"]["final_loss"], d["instrumented_pass"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:18]:
        print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
