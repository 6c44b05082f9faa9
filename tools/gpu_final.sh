#!/usr/bin/env bash
# Final evidence of a state: full GPU suite, smoke, the driver's bench invocations (own arm with all legs, reference arm)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/f_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "reference arm rc=$?"; tail -1 gpurun_out/f_bench_ref.json | cut -c1-600
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "loss", d["config"]["final_loss"])
    print("instrumented", d["instrumented_pass"], "clocks", d["clocks"], "launches", d["gpu_launches"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "share_of_step")})
    print("graphed_step", d.get("graphed_step"))
    print("cpu_baseline", d.get("cpu_baseline"))
    print("torch_eager_gpu", d.get("torch_eager_gpu"))
    print("extras", {k: (v.get("value"), v.get("ms_per_step"), v.get("graphed")) for k, v in d.get("extras", {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
