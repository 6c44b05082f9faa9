#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/k_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/k_suite.log
for ov in 0 1; do
XM_OVERLAP_BRANCHES=$ov timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-extras > gpurun_out/k_bench_ov$ov.json 2> gpurun_out/k_bench_ov$ov.err; echo "bench ov=$ov rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/k_bench_ov$ov.json").read().strip().splitlines()[-1])
    print("overlap $ov: ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "loss", d["config"]["final_loss"], "own", d["own_kernels_ms_per_step"])
    if $ov == 0:
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:16]:
            print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
done
