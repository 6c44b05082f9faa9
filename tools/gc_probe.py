#!/usr/bin/env python
"""Which objects of one training step are only reclaimed by Python's cyclic garbage collector?  (They keep their CUDA
tensors alive until a collection runs, and a generation-2 collection in the middle of a timed region stalls the launch
thread.)  python tools/gc_probe.py"""
import collections
import gc
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer  # noqa: E402

torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = PairedBridgeModel(64, 200, None, 128, 64, 128, 0.3, 0.4, "v4").cuda().train()
tr = PairedTrainer(model)
eeg, roi, _ = synthetic.paired_batch(B, 64, 500, 200, 100, 16, seed=42)
eeg, roi = eeg.cuda(), roi.cuda()
for overlap in (False, True):
    model.overlap_branches = overlap
    tr.step(eeg, roi)
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()
    m0 = torch.cuda.memory_allocated()
    tr.step(eeg, roi)
    torch.cuda.synchronize()
    m1 = torch.cuda.memory_allocated()
    gc.set_debug(gc.DEBUG_SAVEALL)
    n = gc.collect()
    c = collections.Counter(type(o).__name__ for o in gc.garbage)
    tens = [o for o in gc.garbage if torch.is_tensor(o)]
    print(f"overlap={overlap}: {n} unreachable objects, {len(tens)} tensors ({sum(t.numel() * t.element_size() for t in tens) / 2**20:.1f} MiB); "
          f"allocated before / after the step {m0 / 2**20:.0f} / {m1 / 2**20:.0f} MiB")
    print("   types:", c.most_common(12))
    for o in gc.garbage:
        if isinstance(o, dict) and len(o) < 12:
            ks = list(o.keys())[:10]
            print("   dict keys:", ks)
    gc.set_debug(0)
    gc.garbage.clear()
    gc.enable()
