#!/usr/bin/env python
"""Per-step CUDA-event times of the device-resident paired step (bench shape) right after 3 warm-up steps: does the time
of the first timed steps differ from the steady state?  python tools/step_jitter.py [--steps 40]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--gc", choices=["default", "off"], default="default", help="off: gc.collect() + gc.disable() before the timed loop")
ap.add_argument("--smi", type=int, default=0, help="> 0: an `nvidia-smi -lms <ms>` clock query runs next to the loop, as in bench.py")
a = ap.parse_args()
torch.manual_seed(0)
model = PairedBridgeModel(64, 200, None, 128, 64, 128, 0.3, 0.4, "v4").cuda().train()
tr = PairedTrainer(model)
eeg, roi, _ = synthetic.paired_batch(4096, 64, 500, 200, 100, 16, seed=42)
eeg, roi = eeg.cuda(), roi.cuda()
smi = None
if a.smi > 0:
    import subprocess
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                            "--format=csv,noheader,nounits", "-lms", str(a.smi), "-i", "0"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
for _ in range(a.warmup):
    tr.step(eeg, roi)
torch.cuda.synchronize()
import gc  # noqa: E402
import time  # noqa: E402
if a.gc == "off":
    gc.collect()
    gc.disable()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
cpu = []
ev[0].record()
for i in range(a.steps):
    t0 = time.perf_counter()
    tr.step(eeg, roi)
    cpu.append(round((time.perf_counter() - t0) * 1e3, 2))
    ev[i + 1].record()
torch.cuda.synchronize()
if smi is not None:
    smi.terminate()
ms = [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(a.steps)]
print(json.dumps({"smi_ms": a.smi, "slow_steps": [(i, t) for i, t in enumerate(ms) if t > 1.06 * sorted(ms)[len(ms) // 2]], "median": sorted(ms)[len(ms) // 2], "mean": round(sum(ms) / len(ms), 3), "gc": a.gc, "gc_counts": gc.get_count(), "cpu_ms_per_step_call": cpu[:12], "ms_per_step": ms[:12], "first10_mean": round(sum(ms[:10]) / 10, 3), "rest_mean": round(sum(ms[10:]) / max(len(ms) - 10, 1), 3),
                  "mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
                  "alloc_retries": torch.cuda.memory_stats().get("num_alloc_retries"),
                  "segments": torch.cuda.memory_stats().get("segment.all.current")}))
