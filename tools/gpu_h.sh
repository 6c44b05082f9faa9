#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "corrcoef or ffn or tail_fused or d128" > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/h_tests.log
timeout 300 python tools/ffn_bench.py --only-fused > gpurun_out/h_ffn_bench.json 2> gpurun_out/h_ffn_bench.err; cat gpurun_out/h_ffn_bench.json
for f in 0 1 2 3; do
  timeout 120 python tools/ffn_trace.py --flags $f > gpurun_out/h_fwd_trace_f$f.json 2> gpurun_out/h_fwd_trace_f$f.err
  python -c "
import json; d=json.load(open('gpurun_out/h_fwd_trace_f$f.json')); print('fwd flags', $f, 'kernel_ms', round(d['kernel_ms'],3), json.dumps(d['steady_state']))
for r in d['mma'][18:22]: print('   ', r)"
done
timeout 120 python tools/ffn_trace.py --dgrad > gpurun_out/h_dgrad_trace.json 2> gpurun_out/h_dgrad_trace.err
python -c "
import json; d=json.load(open('gpurun_out/h_dgrad_trace.json')); print('dgrad kernel_ms', round(d['kernel_ms'],3), json.dumps(d['steady_state']))
for r in d['mma'][24:36]: print('   ', r)
for r in d['transform_g0'][8:12]: print('   t', r)"
timeout 300 python tools/step_calls.py --filter roi_ > gpurun_out/h_calls.txt 2>&1; cat gpurun_out/h_calls.txt
