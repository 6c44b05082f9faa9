#!/usr/bin/env python
"""Every C-ABI call of ONE paired training step at the bench shape, in launch order, with its CUDA-event time and
algorithmic GB/s -- to find which call of a multi-call entry point is the slow one.
    python tools/step_calls.py [--batch 4096] [--conn device|host] [--filter bn_act]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops, synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--conn", default="device")
    ap.add_argument("--filter", default="")
    a = ap.parse_args()
    torch.manual_seed(42)
    m = PairedBridgeModel(64, 200, 40000, 128, 64, 128, 0.3, 0.4, "v4").cuda().train()
    tr = PairedTrainer(m)
    eeg, roi, conn = (t.cuda() for t in synthetic.paired_batch(a.batch, 64, 500, 200, 100, 40000 if a.conn == "host" else 16, seed=42))
    args = (eeg, roi, conn) if a.conn == "host" else (eeg, roi)
    for _ in range(3):
        tr.step(*args)
    ops.start_timeline()
    tr.step(*args)
    calls = ops.stop_timeline(raw=True)
    tot = 0.0
    for i, (name, ms, fl, by) in enumerate(calls):
        tot += ms
        if a.filter in name:
            print(f"{i:4d} {name:34s} {ms:8.4f} ms  {by / (ms * 1e-3) / 1e9 if ms > 0 else 0:8.1f} GB/s  {fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:7.1f} TF")
    print(f"total {tot:.3f} ms over {len(calls)} calls")


if __name__ == "__main__":
    main()
