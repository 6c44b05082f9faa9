#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/determinism_check.py > gpurun_out/w_det.log 2>&1; cat gpurun_out/w_det.log | tail -40
timeout 300 python tools/step_jitter.py --steps 14 > gpurun_out/w_jit_a.json 2>/dev/null; cat gpurun_out/w_jit_a.json
timeout 300 python tools/step_jitter.py --steps 14 --gc off > gpurun_out/w_jit_b.json 2>/dev/null; cat gpurun_out/w_jit_b.json
timeout 300 python tools/step_jitter.py --steps 14 > gpurun_out/w_jit_c.json 2>/dev/null; cat gpurun_out/w_jit_c.json
timeout 300 python tools/step_jitter.py --steps 14 --gc off > gpurun_out/w_jit_d.json 2>/dev/null; cat gpurun_out/w_jit_d.json
