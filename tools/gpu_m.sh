#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/corr_bench.py > gpurun_out/m_corr.json 2>&1; cat gpurun_out/m_corr.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corrcoef -s 3 -c 1 -o gpurun_out/m_corr python tools/corr_bench.py > gpurun_out/m_ncu.log 2>&1
ncu -i gpurun_out/m_corr.ncu-rep --page raw --csv > gpurun_out/m_corr_raw.csv 2>/dev/null
ncu -i gpurun_out/m_corr.ncu-rep --page source --csv > gpurun_out/m_corr_src.csv 2>/dev/null
XM_PRINT_ERRS=1 timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_paired_step.py -q -s -m gpu 2>&1 | grep "^\[err\]" > gpurun_out/m_errs.txt; wc -l gpurun_out/m_errs.txt
