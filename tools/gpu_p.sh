#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "ffn or tail_fused or d128" > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/p_tests.log
timeout 300 python tools/ffn_bench.py --only-fused > gpurun_out/p_ffn_bench.json 2> gpurun_out/p_ffn_bench.err; cat gpurun_out/p_ffn_bench.json
timeout 120 python tools/ffn_trace.py --dgrad > gpurun_out/p_dgrad_trace.json 2> gpurun_out/p_dgrad_trace.err
python -c "
import json; d=json.load(open('gpurun_out/p_dgrad_trace.json')); print('dgrad kernel_ms', round(d['kernel_ms'],3), json.dumps(d['steady_state']))
for r in d['mma'][24:36]: print('   ', r)
for r in d['transform_g0'][8:12]: print('   t', r)"
