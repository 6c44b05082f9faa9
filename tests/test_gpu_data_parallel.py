"""Data-parallel paired step on real GPUs, collected by `pytest -m gpu`: launches tests/dp_gpu_check.py under torchrun
(one process per GPU, NCCL) and requires every exchange path -- peer-memory in-GEMM reads, peer-memory gather-once,
NCCL all-gather -- to reproduce the single-process global-batch oracle.  Skipped on a 1-GPU box; the host-side logic
of the same step is covered on CPU by tests/test_host_pipeline_cpu.py (world_size-2 gloo)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_step_equals_global_batch_oracle_on_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the node")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(HERE, "dp_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(HERE))
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("dp_gpu_check")]
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert len(lines) >= 3 and all(ln.rstrip().endswith("OK") for ln in lines), "\n".join(lines)


def test_graph_replay_of_the_sharded_step_equals_eager_steps_on_gpus():
    """tests/dp_graph_check.py under torchrun: every rank captures its data-parallel step in a CUDA graph (peer exchanges,
    symmetric-memory barrier, NCCL collectives inside) and three replays equal three eager steps of a twin trainer."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the node")
    if os.environ.get("XM_TEST_DP_GRAPH", "1") == "0":
        pytest.skip("disabled by XM_TEST_DP_GRAPH=0")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29542", os.path.join(HERE, "dp_graph_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=os.path.dirname(HERE))
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("dp_graph_check")]
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert len(lines) >= 2 and all(ln.rstrip().endswith("OK") for ln in lines), "\n".join(lines)
