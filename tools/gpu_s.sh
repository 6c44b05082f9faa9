#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv or split or precise or bn_ or window" > gpurun_out/s_k.log 2>&1; echo "kernel tests rc=$?"; tail -5 gpurun_out/s_k.log
timeout 600 python -m pytest tests/test_gpu_paired_step.py tests/test_gpu_modules.py -x -q -m gpu > gpurun_out/s_step.log 2>&1; echo "step tests rc=$?"; tail -3 gpurun_out/s_step.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-extras > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "loss", d["config"]["final_loss"], d["instrumented_pass"]["ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:22]:
        print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
