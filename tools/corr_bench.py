#!/usr/bin/env python
"""Event-timed connectivity kernel at the bench shape (4096 samples, 100 TR x 200 ROI): python tools/corr_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops  # noqa: E402

x = torch.randn(4096, 100, 200, device="cuda")
res = {}
for prepared in (False, True):
    for _ in range(3):
        ops.roi_corrcoef(x, prepared)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.roi_corrcoef(x, prepared)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = x.numel() * 4 + 4096 * 40000 * 4 * (3 if prepared else 1)
    res["prepared" if prepared else "plain"] = {"ms": round(ms, 4), "gbs": round(by / ms / 1e6, 1), "tflops": round(2 * 4096 * 100 * 40000 / ms / 1e9, 1)}
print(json.dumps(res))
