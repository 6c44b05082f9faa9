"""Eager vs CUDA-graph replay of the paired training step (PairedTrainer.capture), CUDA-event timed.
usage: python tools/graph_bench.py [batch ...]   (default: 256 4096) -> one JSON line per batch"""
import gc
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_eeg_fmri_b200 import ops, synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer  # noqa: E402


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    gc.enable()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda", 0)
    for B in [int(a) for a in sys.argv[1:]] or [256, 4096]:
        torch.manual_seed(42)
        m = PairedBridgeModel(64, 200, 40000, 128, 64, 128, 0.3, 0.4, "v4").to(dev).train()
        tr = PairedTrainer(m)
        eeg, roi, _ = (t.to(dev) for t in synthetic.paired_batch(B, 64, 500, 200, 100, 16, seed=42))
        for _ in range(5):
            tr.step(eeg, roi)
        n = 20 if B <= 1024 else 10
        n0 = ops.launch_count()
        ms_eager = timed(lambda: tr.step(eeg, roi), n)
        calls = (ops.launch_count() - n0) / n
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        g = tr.capture(eeg, roi)
        for _ in range(3):
            g.replay()
        ms_graph = timed(g.replay, n)
        losses = [float(g.replay()) for _ in range(3)]
        print(json.dumps({"batch": B, "ms_eager": round(ms_eager, 3), "ms_graph": round(ms_graph, 3),
                          "c_abi_calls_per_step": calls, "calls_captured": g.launches_captured,
                          "samples_per_s_eager": round(B / ms_eager * 1e3, 1), "samples_per_s_graph": round(B / ms_graph * 1e3, 1),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2), "losses": losses}), flush=True)
        del g, tr, m
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
