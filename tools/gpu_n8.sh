#!/usr/bin/env bash
# N-GPU validation: data-parallel parity check on all GPUs, then the weak-scaling bench line (N from $1, default 8)
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dp_gpu_check.py > gpurun_out/n${N}_dp.log 2>&1; echo "dp rc=$?"; grep dp_gpu_check gpurun_out/n${N}_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/dp_graph_check.py > gpurun_out/n${N}_dpg.log 2>&1; echo "dp graph rc=$?"; grep dp_graph_check gpurun_out/n${N}_dpg.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 8 --warmup 3 --no-eager > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n${N}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/n${N}_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "loss", d["config"]["final_loss"])
    print("strong", d.get("strong_scaling"))
    print("graphed", d.get("graphed_step"))
    print("instrumented", d.get("instrumented_pass"), d.get("clocks"))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
        if "infonce" in k or "peer" in k: print("   ", k, v["calls_per_step"], v["ms_per_step"])
except Exception as e:
    print("parse failed", e)
PY
