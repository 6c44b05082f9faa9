#!/usr/bin/env bash
# GPU session A (round 2): fused FFN + general attention kernel tests, the full suite, bench A/B of the fused FFN.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ffn or general_attention or transformer_block_masks or tail_fused" > gpurun_out/a_new_tests.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/a_new_tests.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/a_suite.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err; echo "bench rc=$?"
XM_FUSED_FFN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/a_bench_unfused.json 2> gpurun_out/a_bench_unfused.err
python - <<'PY'
import json
for n in ("fused", "unfused"):
    try:
        d = json.loads(open(f"gpurun_out/a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["final_loss"])
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:12]:
            print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
    except Exception as e:
        print(n, "failed", e)
PY
