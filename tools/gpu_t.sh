#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/t_k.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/t_k.log
timeout 300 python tools/attn_bias_ab.py > gpurun_out/t_attn_ab.json 2> gpurun_out/t_attn_ab.err; cat gpurun_out/t_attn_ab.json; tail -2 gpurun_out/t_attn_ab.err
