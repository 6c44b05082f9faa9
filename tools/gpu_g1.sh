#!/usr/bin/env bash
# session: CUDA-graph step -- new tests, then eager vs replay timings
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphed_step.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/g1_tests.log
timeout 600 python tools/graph_bench.py 256 4096 2>&1 | tail -20 | tee gpurun_out/g1_graph_bench.log
