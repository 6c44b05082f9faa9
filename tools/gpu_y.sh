#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/determinism_check.py 2>&1 | grep -v Warning | tail -6 > gpurun_out/y_det.log; cat gpurun_out/y_det.log
for i in 1 2 3; do
XM_PRINT_ERRS=1 timeout 600 python -m pytest tests/test_gpu_paired_step.py -x -q -m gpu -s -k "baseline_shape_parity" 2>&1 | grep -E "err\] grad|passed|failed" | sort -t' ' -k5 -g -r | head -3
done > gpurun_out/y_parity.log 2>&1; cat gpurun_out/y_parity.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "infonce" > gpurun_out/y_nce.log 2>&1; echo "nce tests rc=$?"; tail -2 gpurun_out/y_nce.log
