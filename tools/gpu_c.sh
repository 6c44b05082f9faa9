#!/usr/bin/env bash
# GPU session C: FFN v2 kernels (tests, event timing, ncu), full-scale parity with the 3-pass first convs, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ffn or tail_fused" > gpurun_out/c_ffn_tests.log 2>&1; echo "ffn tests rc=$?"; tail -4 gpurun_out/c_ffn_tests.log
timeout 300 python tools/ffn_bench.py > gpurun_out/c_ffn_bench.json 2> gpurun_out/c_ffn_bench.err; cat gpurun_out/c_ffn_bench.json
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/c_suite.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
XM_CONV_PRECISE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c_bench_convsingle.json 2> gpurun_out/c_bench_convsingle.err
python - <<'PY'
import json
for n in ("c_bench", "c_bench_convsingle"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["final_loss"])
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]:
            print("   ", k, v["calls_per_step"], v["ms_per_step"], v["tflops"], v["gbs"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 python tools/full_scale_parity.py --batch 2048 --fp64 --out gpurun_out/c_full_scale_parity.json > gpurun_out/c_parity.log 2>&1; echo "parity rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c_full_scale_parity.json"))
    v = d["vs_fp64"]
    print("loss rel err", v["loss_rel_err_gpu"], "grad median", v["grad_rel_err_gpu_median"], "max", v["grad_rel_err_gpu_max"])
    print({k: round(x["gpu"], 6) for k, x in v["worst_gpu"].items()})
except Exception as e:
    print("parity failed", e)
PY
for k in ffn_fwd ffn_dgrad; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/c_$k python tools/ffn_bench.py --only-fused --reps 1 > gpurun_out/c_ncu_$k.log 2>&1
  ncu -i gpurun_out/c_$k.ncu-rep --page raw --csv > gpurun_out/c_${k}_raw.csv 2>/dev/null
done
