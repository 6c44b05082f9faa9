"""Drop-in for the hot-path part of EEG_CODE/eeg_data_utils.py plus the device preprocessing the
north-star attributes to this module: window index generation, window gather and band power
(no reference implementation: SURVEY.md section 0; definitions in oracle/spectral.py), and
`normalize_modality` (EEG_CODE/run_training_lite.py:48-51).  The .mat / HDF5 readers are out of scope."""
from __future__ import annotations

import math
import os
from fractions import Fraction
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops

__all__ = ["load_eeg_labels", "window_indices", "gather_windows", "band_power", "band_bins", "normalize_modality",
           "DEFAULT_BANDS"]

# EEG_CODE/config.py:35 names the bands; half-open [lo, hi) Hz
DEFAULT_BANDS: Dict[str, Tuple[float, float]] = {"theta": (4.0, 8.0), "alpha": (8.0, 13.0), "beta": (13.0, 30.0)}


def load_eeg_labels(label_dir, binary: bool = True) -> Dict[int, int]:
    """EEG_CODE/eeg_data_utils.py:19-43: labels from <label_dir>/medical_score.csv; 'subNN' -> NN;
    label = 0 if score <= 2 else (1 if binary else score).  Host side (runs once)."""
    import pandas as pd

    csv_path = os.path.join(str(label_dir), "medical_score.csv")
    if not os.path.exists(csv_path):
        raise FileNotFoundError(f"Label file not found: {csv_path}")
    df = pd.read_csv(csv_path).dropna(subset=["Postoperative evaluation"])
    subj = df["Subject"]
    # the reference tests `dtype == object`; newer pandas types text columns as `str`
    is_text = subj.dtype == object or pd.api.types.is_string_dtype(subj)
    ids = subj.str.replace("sub", "", regex=False).astype(int) if is_text else subj.astype(int)
    out = {}
    for sid, score in zip(ids.tolist(), df["Postoperative evaluation"].tolist()):
        out[int(sid)] = 0 if score <= 2 else 1 if binary else score
    return out


def window_indices(n_rec: int, n_samples: int, win: int, hop: int, rec_labels: Optional[torch.Tensor] = None,
                   rec_subjects: Optional[torch.Tensor] = None, device="cuda"):
    """int64 device tensors (starts, rec_ids, labels, subjects) of the n_rec * ((n-win)//hop + 1) windows;
    windows never cross recordings."""
    if n_samples < win or n_rec == 0:
        z = torch.empty(0, dtype=torch.int64, device=device)
        return z, z.clone(), (None if rec_labels is None else z.clone()), (None if rec_subjects is None else z.clone())
    as64 = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.int64, device=device).contiguous()
    return ops.window_index(n_rec, n_samples, win, hop, as64(rec_labels), as64(rec_subjects), device=device)


def gather_windows(rec: torch.Tensor, win: int, hop: int, channels_last: bool = False) -> torch.Tensor:
    """rec (R, C, n) -> (R*n_win, C, win), or (R*n_win, win, C) with channels_last=True."""
    if rec.shape[0] == 0 or rec.shape[1] == 0 or rec.shape[2] < win:  # no complete window
        shape = (0, win, rec.shape[1]) if channels_last else (0, rec.shape[1], win)
        return torch.empty(shape, device=rec.device, dtype=torch.float32)
    return ops.window_gather(rec, win, hop, channels_last=channels_last)


def band_bins(bands: Sequence[Tuple[float, float]], nfft: int, fs: float):
    """Half-open integer bin ranges [k_lo, k_hi) with lo <= k*fs/nfft < hi (exact rational arithmetic)."""
    fsr = Fraction(fs).limit_denominator(1_000_000)
    out = []
    for lo, hi in bands:
        k_lo = max(math.ceil(Fraction(lo).limit_denominator(1_000_000) * nfft / fsr), 0)
        k_hi = min(math.ceil(Fraction(hi).limit_denominator(1_000_000) * nfft / fsr), nfft // 2 + 1)
        out += [k_lo, max(k_hi, k_lo)]
    return out


_taper_cache: Dict[Tuple[int, str], Tuple[torch.Tensor, float]] = {}


def _hann(win: int, device) -> Tuple[torch.Tensor, float]:
    key = (win, str(device))
    if key not in _taper_cache:
        t = torch.hann_window(win, periodic=True, dtype=torch.float64).to(torch.float32)
        _taper_cache[key] = (t.to(device).contiguous(), float((t.double() ** 2).sum()))
    return _taper_cache[key]


def band_power(rec: torch.Tensor, fs: float, win: int, hop: int, bands=None, nfft: Optional[int] = None,
               path: str = "auto") -> torch.Tensor:
    """Band power of every window of rec (R, C, n): Hann taper, rfft zero-padded to nfft (next power of
    two >= win by default), one-sided PSD, sum over [lo, hi) bins times the bin width.
    Returns (R*n_win, C, n_bands) fp32; windows are read in place (never materialised).  path: "auto" picks the
    tensor-core DFT kernel when only a few bins are needed (<= 32) and there are >= 64 channels, else the FFT
    kernel; "fft" / "dft" force one."""
    bands = list((bands or DEFAULT_BANDS).values()) if isinstance(bands or DEFAULT_BANDS, dict) else list(bands)
    if rec.dim() != 3:
        raise ValueError("rec must be (recordings, channels, samples)")
    if rec.shape[0] == 0 or rec.shape[1] == 0 or rec.shape[2] < win:  # no complete window: empty result
        return torch.empty(0, rec.shape[1], len(bands), device=rec.device, dtype=torch.float32)
    if nfft is None:
        nfft = 1 << max(6, (win - 1).bit_length())
    taper, sumsq = _hann(win, rec.device)
    host_bins = band_bins(bands, nfft, fs)
    bins = torch.tensor(host_bins, dtype=torch.int32, device=rec.device)
    total = sum(host_bins[2 * b + 1] - host_bins[2 * b] for b in range(len(bands)))
    return ops.bandpower(rec, win, hop, nfft, fs, taper, sumsq, bins, total_bins=total, path=path)


def normalize_modality(feat: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """EEG_CODE/run_training_lite.py:48-51: (x - mean) / (std + eps) over the whole tensor, population std."""
    return ops.zscore(feat.reshape(1, -1), eps).reshape(feat.shape)
