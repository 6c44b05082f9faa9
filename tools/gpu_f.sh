#!/usr/bin/env bash
# GPU session F: MMA issue-rate probe; full suite with the fused optimizer / derived connectivity; new bench line
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/mma_rate tools/probes/mma_rate.cu > gpurun_out/f_probe_build.log 2>&1 && timeout 60 /tmp/mma_rate > gpurun_out/f_mma_rate.txt 2>&1; cat gpurun_out/f_mma_rate.txt
grep MemTotal /proc/meminfo; nproc
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/f_suite.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "loss", d["config"]["final_loss"], "launches", d["gpu_launches"])
    print("step_roofline", d["step_roofline"])
    print("extras", json.dumps(d.get("extras"), indent=1))
    print("eager", d.get("torch_eager_gpu"))
    print("cpu", d.get("cpu_baseline"))
    print("torch ops ms", d["torch_ops_ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:16]:
        print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/f_ref.json; tail -4 gpurun_out/f_ref.err
