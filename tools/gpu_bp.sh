#!/usr/bin/env bash
# session: band-power overlap kernel vs v1 -- parity tests, wait-cycle trace, config-5 sweep
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bandpower" 2>&1 | tail -15 | tee gpurun_out/bp_tests.log
timeout 300 python tools/bp_trace.py 2>&1 | tail -1 | tee gpurun_out/bp_trace.log
for k in 2 1 2; do
  echo "kernel $k"
  XM_BP_DFT_KERNEL=$k timeout 300 python tools/bandpower_sweep.py --windows 131072 --paths dft 2>&1 | tail -1 | cut -c1-330
done | tee gpurun_out/bp_sweep.log
