#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "cross2" > gpurun_out/z_k.log 2>&1; echo "cross2 tests rc=$?"; tail -3 gpurun_out/z_k.log
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_paired_step.py -x -q -m gpu > gpurun_out/z_m.log 2>&1; echo "module + step tests rc=$?"; tail -3 gpurun_out/z_m.log
timeout 600 python -m pytest tests -x -q -m gpu -k "xai or bridge or saliency or attribution" > gpurun_out/z_b.log 2>&1; echo "bridge tests rc=$?"; tail -2 gpurun_out/z_b.log
