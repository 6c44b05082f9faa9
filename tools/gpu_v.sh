#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
XM_PRINT_ERRS=1 timeout 600 python -m pytest tests/test_gpu_paired_step.py -x -q -m gpu -s -k "baseline_shape_parity" 2>&1 | grep -E "err\] grad|passed|failed" | sort -t' ' -k5 -g -r | head -8
done > gpurun_out/v_parity.log 2>&1; cat gpurun_out/v_parity.log
timeout 300 python tools/step_jitter.py > gpurun_out/v_jitter.json 2> gpurun_out/v_jitter.err; cat gpurun_out/v_jitter.json; tail -2 gpurun_out/v_jitter.err
timeout 300 python tools/step_jitter.py > gpurun_out/v_jitter2.json 2> gpurun_out/v_jitter2.err; cat gpurun_out/v_jitter2.json
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/v_conv.log 2>&1; echo "conv tests rc=$?"; tail -3 gpurun_out/v_conv.log
