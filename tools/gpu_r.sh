#!/usr/bin/env bash
# fused InfoNCE backward: kernel test (under its own timeout: a protocol error would hang), step parity, bench A/B
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "infonce" > gpurun_out/r_nce.log 2>&1; echo "nce tests rc=$?"; tail -15 gpurun_out/r_nce.log
timeout 600 python -m pytest tests/test_gpu_paired_step.py tests/test_gpu_modules.py -x -q -m gpu > gpurun_out/r_step.log 2>&1; echo "step tests rc=$?"; tail -3 gpurun_out/r_step.log
for f in 1 0; do
XM_FUSED_INFONCE_BWD=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-extras > gpurun_out/r_bench_$f.json 2> gpurun_out/r_bench_$f.err; echo "bench fused=$f rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r_bench_$f.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "loss", d["config"]["final_loss"], d["instrumented_pass"]["ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
        if "infonce" in k or "split3" in k: print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
done
