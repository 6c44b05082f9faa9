#!/usr/bin/env bash
# final evidence + ncu --set full of the band-power overlap kernel (after its own command has exited 0 without ncu)
set -u
bash tools/gpu_final.sh
timeout 300 python tools/bandpower_sweep.py --windows 65536 --paths dft > gpurun_out/n_bp_small.json 2>&1; echo "bp rc=$?"; tail -1 gpurun_out/n_bp_small.json | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bandpower_dft2 -c 2 -o gpurun_out/n_bp2 python tools/bandpower_sweep.py --windows 65536 --paths dft > gpurun_out/n_ncu_bp2.log 2>&1; echo "ncu bp rc=$?"
ncu -i gpurun_out/n_bp2.ncu-rep --page raw --csv > gpurun_out/n_bp2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/n_bp2_raw.csv gpurun_out/n_bp2.md | head -5
