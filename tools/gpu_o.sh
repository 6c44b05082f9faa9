#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for f in 0 1 4 5; do
  timeout 120 python tools/ffn_trace.py --dgrad --flags $f > gpurun_out/o_dgrad_f$f.json 2> gpurun_out/o_dgrad_f$f.err
  python -c "
import json; d=json.load(open('gpurun_out/o_dgrad_f$f.json')); print('dgrad flags', $f, 'kernel_ms', round(d['kernel_ms'],3), json.dumps(d['steady_state']))
for r in d['mma'][24:30]: print('   ', r)"
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/o_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/o_suite.log
