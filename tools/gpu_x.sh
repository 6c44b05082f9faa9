#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/gc_probe.py 256 > gpurun_out/x_gc.log 2>&1; grep -v Warning gpurun_out/x_gc.log | head -30
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/x_conv.log 2>&1; echo "conv tests rc=$?"; tail -3 gpurun_out/x_conv.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-extras > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/x_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "loss", d["config"]["final_loss"], d["instrumented_pass"]["ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
        if "conv" in k or "bn_" in k: print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
