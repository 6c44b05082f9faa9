"""Max relative error of the DFT band-power kernel against the fp64 oracle at the config-5 shape (tones + noise)."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_eeg_fmri_b200 import eeg_data_utils as edu  # noqa: E402
from oracle import spectral as osp  # noqa: E402

torch.manual_seed(8)
R, C, n, win, hop, fs = 2, 128, 1024 + 20 * 512, 1024, 512, 1000.0
t = torch.arange(n, dtype=torch.float64) / fs
for name, amp in (("noise+tones", 1.0), ("strong tones", 50.0)):
    tones = sum(a * amp * torch.sin(2 * math.pi * f * t + ph) for a, f, ph in ((2.0, 6.0, 0.3), (1.5, 10.0, 1.1), (1.0, 20.0, 2.0)))
    rec = (torch.randn(R, C, n, dtype=torch.float64) + tones).float().cuda()
    p = edu.band_power(rec, fs, win, hop, path="dft").cpu().numpy().astype(np.float64)
    want = osp.gather_windows(rec.cpu().numpy(), win, hop)
    ref = osp.band_power(want, fs, nfft=1024, taper=torch.hann_window(win, periodic=True, dtype=torch.float64).float().double().numpy())
    err = np.abs(p - ref) / np.maximum(ref, 1e-30)
    signed = ((p - ref) / np.maximum(ref, 1e-30)).mean()
    print(f"{name}: max rel err {err.max():.3e}  mean signed {signed:+.3e}  kernel={os.environ.get('XM_BP_DFT_KERNEL', '2')} chunk={os.environ.get('XM_BP_CHUNK', '8')}")
