#!/usr/bin/env bash
# Evidence session: event-timed micro-benches, then ncu --set full of the new kernels (each only after its own command
# has exited 0 without ncu), then the launch list of the bench command.
set -u
mkdir -p gpurun_out
timeout 300 python tools/nce_bench.py > gpurun_out/n_nce_bench.json 2> gpurun_out/n_nce_bench.err; echo "nce_bench rc=$?"; cat gpurun_out/n_nce_bench.json
timeout 300 python tools/ffn_bench.py > gpurun_out/n_ffn_bench.json 2> gpurun_out/n_ffn_bench.err; echo "ffn_bench rc=$?"; cat gpurun_out/n_ffn_bench.json
timeout 300 python tools/corr_bench.py > gpurun_out/n_corr_bench.json 2>/dev/null; cat gpurun_out/n_corr_bench.json
timeout 300 python tools/bandpower_sweep.py --windows 65536 > gpurun_out/n_bp_small.json 2>&1; tail -1 gpurun_out/n_bp_small.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:nce_kernel -c 4 -o gpurun_out/n_nce python tools/nce_bench.py --only-fused --reps 1 > gpurun_out/n_ncu_nce.log 2>&1; echo "ncu nce rc=$?"
timeout 900 $NCU -k regex:ffn_ -c 4 -o gpurun_out/n_ffn python tools/ffn_bench.py --only-fused --reps 1 > gpurun_out/n_ncu_ffn.log 2>&1; echo "ncu ffn rc=$?"
timeout 600 $NCU -k regex:corrcoef -c 1 -o gpurun_out/n_corr python tools/corr_bench.py > gpurun_out/n_ncu_corr.log 2>&1; echo "ncu corr rc=$?"
timeout 600 $NCU -k regex:bandpower_dft -c 2 -o gpurun_out/n_bp python tools/bandpower_sweep.py --windows 65536 > gpurun_out/n_ncu_bp.log 2>&1; echo "ncu bp rc=$?"
for n in nce ffn corr bp; do
  ncu -i gpurun_out/n_$n.ncu-rep --page raw --csv > gpurun_out/n_${n}_raw.csv 2>/dev/null
done
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu --no-eager --no-extras > gpurun_out/n_bench_plain.json 2> gpurun_out/n_bench_plain.err; echo "bench rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 -c 700 --csv --log-file gpurun_out/n_launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu --no-eager --no-extras > gpurun_out/n_ncu_bench.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/n_launches.csv gpurun_out/n_launches.md | head -40
