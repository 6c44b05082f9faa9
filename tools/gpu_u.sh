#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "bandpower or window or spectral or smoke or preprocess" > gpurun_out/u_bp.log 2>&1; echo "bandpower tests rc=$?"; tail -3 gpurun_out/u_bp.log
timeout 300 python tools/bandpower_sweep.py --windows 131072 > gpurun_out/u_bp_sweep.json 2>&1; tail -2 gpurun_out/u_bp_sweep.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
