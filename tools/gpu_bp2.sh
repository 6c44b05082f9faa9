#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k bandpower 2>&1 | tail -3
timeout 300 python tools/bp_err.py 2>&1 | tail -2
for k in 2 1; do echo "kernel $k"; XM_BP_DFT_KERNEL=$k timeout 300 python tools/bandpower_sweep.py --windows 1048576 --paths dft 2>&1 | tail -1 | cut -c1-330; done
} | tee gpurun_out/bp_final.log
