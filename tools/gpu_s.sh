#!/usr/bin/env bash
# repeatability of the bench line: the same short invocation six times
set -u
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-extras 2>/dev/null | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['instrumented_pass']['ms_per_step'], d['clocks'])"
done | tee gpurun_out/s_repeat.log
