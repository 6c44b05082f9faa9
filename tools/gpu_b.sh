#!/usr/bin/env bash
# GPU session B: event-timed fused FFN kernels + ncu --set full of both
set -u
mkdir -p gpurun_out
timeout 300 python tools/ffn_bench.py > gpurun_out/b_ffn_bench.json 2> gpurun_out/b_ffn_bench.err; cat gpurun_out/b_ffn_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ffn_ -c 4 -o gpurun_out/b_ffn python tools/ffn_bench.py --only-fused --reps 1 > gpurun_out/b_ncu.log 2>&1
ncu -i gpurun_out/b_ffn.ncu-rep --page raw --csv > gpurun_out/b_ffn_raw.csv 2>/dev/null
ncu -i gpurun_out/b_ffn.ncu-rep --page details > gpurun_out/b_ffn_details.txt 2>/dev/null
grep -E "ffn_|Duration|Throughput|Pipe|Issue|Stall|L2|DRAM" gpurun_out/b_ffn_details.txt | head -80
