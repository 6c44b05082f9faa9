#!/usr/bin/env bash
# final evidence (suite, smoke, bench, reference arm) + ncu launch list of the bench command (after it has exited 0 without ncu)
set -u
bash tools/gpu_final.sh
timeout 900 python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu --no-eager --no-extras --no-graph > gpurun_out/n_bench_plain.json 2> gpurun_out/n_bench_plain.err; echo "plain bench rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 -c 700 --csv --log-file gpurun_out/n_launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu --no-eager --no-extras --no-graph > gpurun_out/n_ncu_bench.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/n_launches.csv gpurun_out/n_launches.md | head -30
