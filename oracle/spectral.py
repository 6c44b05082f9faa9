"""Authored oracle (test infrastructure): EEG windowing, band power, label / normalisation rules.

PARITY UNPINNED for windowing and band power: the reference contains no implementation
(EEG_CODE/eeg_data_utils.py:46-186 only loads pre-computed MATLAB spectra).  The definitions
below follow SURVEY.md section 8a rows 2-3 and the reference's vocabulary
(EEG_CODE/config.py:34-36 bands; EEG_CODE/CrossModal_EEG_scr.ipynb cell 7:45-47 PW layout) and are
pinned by tests/test_oracle_spectral.py (pure tone, Parseval, scipy.signal.periodogram).

Rules that DO come from the reference:
* `binarise_score`  -- EEG_CODE/eeg_data_utils.py:42 (and the `<=1` variant of
  EEG_CODE/run_training_lite.py:290-291)
* `normalize_modality` -- EEG_CODE/run_training_lite.py:48-51
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Dict, Sequence, Tuple

import numpy as np

# EEG_CODE/config.py:35 names the bands; edges are the conventional ones, half-open [lo, hi).
DEFAULT_BANDS: Dict[str, Tuple[float, float]] = {"theta": (4.0, 8.0), "alpha": (8.0, 13.0), "beta": (13.0, 30.0)}


def n_windows(n_samples: int, win: int, hop: int) -> int:
    return 0 if n_samples < win else (n_samples - win) // hop + 1


def window_indices(n_rec: int, n_samples: int, win: int, hop: int, rec_labels=None, rec_subjects=None):
    """int64 arrays (starts, rec_ids, labels, subjects) for windows that never cross recordings.

    Window g = r * n_win + w covers samples [w*hop, w*hop + win) of recording r.
    """
    nw = n_windows(n_samples, win, hop)
    g = np.arange(n_rec * nw, dtype=np.int64)
    rec_ids = g // max(nw, 1)
    starts = (g - rec_ids * nw) * hop
    labels = None if rec_labels is None else np.asarray(rec_labels, dtype=np.int64)[rec_ids]
    subjects = None if rec_subjects is None else np.asarray(rec_subjects, dtype=np.int64)[rec_ids]
    return starts, rec_ids, labels, subjects


def gather_windows(rec: np.ndarray, win: int, hop: int) -> np.ndarray:
    """rec (R, C, n) -> (R*n_win, C, win)"""
    R, C, n = rec.shape
    nw = n_windows(n, win, hop)
    out = np.empty((R * nw, C, win), dtype=rec.dtype)
    for r in range(R):
        for w in range(nw):
            out[r * nw + w] = rec[r, :, w * hop:w * hop + win]
    return out


def hann_periodic(win: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win) / torch.hann_window(win): periodic Hann, float64."""
    n = np.arange(win, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win)


def band_bins(bands: Sequence[Tuple[float, float]], nfft: int, fs: float) -> np.ndarray:
    """Half-open integer bin ranges [k_lo, k_hi) with k*fs/nfft in [lo, hi), exact rational arithmetic."""
    out = []
    fsr = Fraction(fs).limit_denominator(1_000_000)
    for lo, hi in bands:
        k_lo = math.ceil(Fraction(lo).limit_denominator(1_000_000) * nfft / fsr)
        k_hi = math.ceil(Fraction(hi).limit_denominator(1_000_000) * nfft / fsr)
        k_hi = min(k_hi, nfft // 2 + 1)
        out += [max(k_lo, 0), max(k_hi, k_lo)]
    return np.asarray(out, dtype=np.int32)


def band_power(windows: np.ndarray, fs: float, bands=None, nfft: int | None = None, taper: np.ndarray | None = None):
    """windows (..., win) -> (..., n_bands) float64.

    X = rfft(taper * x, nfft);  one-sided PSD  P_k = |X_k|^2 * s_k / (fs * sum(taper^2)),
    s_k = 2 except DC / Nyquist;  band power = sum_{k in band} P_k * fs / nfft.
    No detrending (scipy.signal.periodogram(..., detrend=False, scaling='density') convention).
    """
    bands = list((bands or DEFAULT_BANDS).values()) if isinstance(bands or DEFAULT_BANDS, dict) else list(bands)
    win = windows.shape[-1]
    nfft = nfft or win
    taper = hann_periodic(win) if taper is None else np.asarray(taper, dtype=np.float64)
    x = windows.astype(np.float64) * taper
    X = np.fft.rfft(x, n=nfft, axis=-1)
    P = (X.real ** 2 + X.imag ** 2)
    s = np.full(nfft // 2 + 1, 2.0)
    s[0] = 1.0
    if nfft % 2 == 0:
        s[-1] = 1.0
    P = P * s / (fs * np.sum(taper ** 2))
    bins = band_bins(bands, nfft, fs)
    out = np.stack([P[..., bins[2 * i]:bins[2 * i + 1]].sum(-1) for i in range(len(bands))], axis=-1)
    return out * (fs / nfft)


def pw_layout(power: np.ndarray) -> np.ndarray:
    """(n_win, C, F) band powers of one recording -> reference PW tensor (C*F, T=n_win), row = c*F + f
    (EEG_CODE/CrossModal_EEG_scr.ipynb cell 7:45-46)."""
    T, C, F = power.shape
    return power.transpose(1, 2, 0).reshape(C * F, T)


def normalize_modality(feat: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """EEG_CODE/run_training_lite.py:48-51 -- global z-score, population std."""
    mean = feat.mean()
    std = feat.std() + eps
    return (feat - mean) / std


def binarise_score(score, binary: bool = True, threshold: float = 2):
    """EEG_CODE/eeg_data_utils.py:42: `0 if score <= 2 else 1 if binary else score`
    (conditional-expression precedence: scores <= threshold map to 0 even when binary=False)."""
    return 0 if score <= threshold else 1 if binary else score
