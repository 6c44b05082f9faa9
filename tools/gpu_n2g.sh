#!/usr/bin/env bash
# 2-GPU: CUDA-graph replay of the data-parallel step
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dp_graph_check.py > gpurun_out/n2g.log 2>&1; echo "dp graph rc=$?"; grep -v "^\*\|Warning\|warn" gpurun_out/n2g.log | tail -25
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
