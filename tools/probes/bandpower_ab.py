import sys, os, json
sys.path.insert(0, os.getcwd())
from pathlib import Path
from multimodal_eeg_fmri_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = Path(sys.argv[1]).resolve()
import torch
from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
R, C, win, hop, wpr = 256, 128, 1024, 512, 64
n = win + (wpr - 1) * hop
g = torch.Generator(device="cuda").manual_seed(1)
bufs = [torch.randn(R, C, n, device="cuda", generator=g) for _ in range(2)]
out = {}
for path in ("dft", "fft"):
    for i in range(3):
        p = edu.band_power(bufs[i & 1], 1000.0, win, hop, path=path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        p = edu.band_power(bufs[i & 1], 1000.0, win, hop, path=path)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    nwin = R * wpr
    out[path] = {"ms": round(ms, 3), "windows_per_s": round(nwin / ms * 1e3), "unique_GBs": round(R * C * n * 4 / ms / 1e6, 1),
                 "checksum": float(p.double().sum())}
print(json.dumps({"lib": str(_lib.LIB_PATH.name), **out}))
