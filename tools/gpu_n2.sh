#!/usr/bin/env bash
# 2-GPU sanity: data-parallel parity on real GPUs + weak/strong scaling bench line
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_data_parallel.py -x -q -m gpu > gpurun_out/n2_dp.log 2>&1; echo "dp rc=$?"; tail -5 gpurun_out/n2_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n2_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/n2_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "loss", d["config"]["final_loss"])
    print("strong", d.get("strong_scaling"))
    print("graphed", d.get("graphed_step"))
    print("extras", {k: v.get("value") for k, v in d.get("extras", {}).items()})
except Exception as e:
    print("parse failed", e)
PY
