#!/usr/bin/env bash
# First GPU call of the next session: everything that was written after the round-1 GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh'
# Outputs land in gpurun_out/ (each step under its own `timeout`, so a hung probe cannot hold the box).
set -u
mkdir -p gpurun_out
NVCC="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo"

# 1. fused FFN probes (never run so far): small ragged shape first, then the bench shape
for k in ffn_fused_fwd ffn_fused_dgrad; do
  $NVCC -o tools/probes/$k tools/probes/$k.cu > gpurun_out/$k.build.log 2>&1 || { echo "$k: build failed"; continue; }
  for args in "1000 0.0" "1000 0.1" "70001 0.1" "1024000 0.1"; do
    echo "== $k $args"; timeout 60 tools/probes/$k $args
  done > gpurun_out/$k.log 2>&1
  tail -4 gpurun_out/$k.log
done

# 2. measurements of the rows widened in round 1 (attribution, shard -> device)
timeout 300 python tools/widen_bench.py > gpurun_out/widen_bench.json 2> gpurun_out/widen_bench.err; cat gpurun_out/widen_bench.json

# 3. the ncu capture the band-power DFT kernel still lacks (after the un-profiled command has exited 0)
if timeout 120 python tools/bandpower_sweep.py --windows 65536 > gpurun_out/bp_small.json 2>&1; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bandpower_dft -c 2 -o gpurun_out/bp_dft \
    python tools/bandpower_sweep.py --windows 65536 > gpurun_out/ncu_bp.log 2>&1
  ncu -i gpurun_out/bp_dft.ncu-rep --page raw --csv > gpurun_out/bp_dft_raw.csv 2>/dev/null
fi
