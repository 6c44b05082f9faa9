"""Drop-in for the hot-path part of EEG_CODE/eeg_data_utils.py plus the device preprocessing the
north-star attributes to this module: window index generation, window gather and band power
(no reference implementation: SURVEY.md section 0; definitions in oracle/spectral.py), and
`normalize_modality` (EEG_CODE/run_training_lite.py:48-51), and the .mat feature readers (:46-186; SURVEY.md
section 8f rank 3: host-side byte movement, bit-exact; MATLAB v7.3 / HDF5 files need `h5py`, which is optional)."""
from __future__ import annotations

import glob
import logging
import math
import os
from fractions import Fraction
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops

logger = logging.getLogger(__name__)

__all__ = ["load_eeg_labels", "load_eeg_conn_features", "load_eeg_pw_features", "load_eeg_erp_features",
           "window_indices", "gather_windows", "band_power", "band_bins", "normalize_modality", "DEFAULT_BANDS"]

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


# ------------------------------------------------------------------------- .mat feature readers (:46-186)
def _first_mat_variable(path, flatten: bool):
    """The first non-private variable of a MATLAB v5 file as fp32 with NaN -> 0 (None if the file has none)."""
    import numpy as np
    from scipy.io import loadmat

    mat = loadmat(path)
    name = next((k for k in mat if not k.startswith("_")), None)
    if name is None:
        return None
    data = np.array(mat[name], dtype=np.float32)
    return np.nan_to_num(data.flatten() if flatten else data, nan=0.0)


def _hdf5_erp(path):
    """ERP array of a MATLAB v7.3 (HDF5) file: group `erp_struct` | `erp` | first key; dataset `avg`, else `trial`
    (3-D trials averaged over axis 0), else the first dataset with >= 2 dims (:141-165).  Returns (found, data);
    raises when the file is not HDF5 or h5py is not installed (the caller then tries the v5 reader)."""
    import h5py  # optional dependency
    import numpy as np

    with h5py.File(path, "r") as hf:
        group = hf["erp_struct"] if "erp_struct" in hf else hf["erp"] if "erp" in hf else hf[list(hf.keys())[0]]
        if "avg" in group:
            data = np.array(group["avg"], dtype=np.float32)
        elif "trial" in group:
            data = np.array(group["trial"], dtype=np.float32)
            if data.ndim == 3:
                data = np.mean(data, axis=0)
        else:
            cand = next((group[k] for k in group.keys() if hasattr(group[k], "shape") and len(group[k].shape) >= 2), None)
            if cand is None:
                return False, None
            data = np.array(cand, dtype=np.float32)
        return True, np.nan_to_num(data, nan=0.0)


def load_eeg_conn_features(conn_dir, subject_list, band_list, cond_list):
    """:46-86 -- {(subject, band_key, condition, 0): flat fp32} from `conn_<BandName>_<cond>_subNN.mat`, falling
    back to `conn_<band_key>_...` when the capitalised name matches nothing; `band_list` maps key -> name."""
    conn_dir = Path(conn_dir)
    out = {}
    for subj in subject_list:
        tag = f"{subj:02d}"
        for band_key, band_name in band_list.items():
            for cond in cond_list:
                files = sorted(glob.glob(str(conn_dir / f"conn_{band_name}_{cond}_sub{tag}.mat")))
                if not files:
                    files = sorted(glob.glob(str(conn_dir / f"conn_{band_key}_{cond}_sub{tag}.mat")))
                for f in files:
                    try:
                        data = _first_mat_variable(f, flatten=True)
                        if data is not None:
                            out[(subj, band_key, cond, 0)] = data
                    except Exception as e:  # noqa: BLE001 - reference behaviour: log and continue
                        logger.warning(f"Error loading {f}: {e}")
    logger.info(f"Loaded {len(out)} EEG connectivity samples")
    return out


def load_eeg_pw_features(pw_dir, subject_list, band_list, freq_list):
    """:89-125 -- {(subject, band, freq, 0): flat fp32} from `powspctrm_<band>_<freq>_subNN.mat`."""
    pw_dir = Path(pw_dir)
    out = {}
    for subj in subject_list:
        tag = f"{subj:02d}"
        for band in band_list:
            for freq in freq_list:
                for f in sorted(glob.glob(str(pw_dir / f"powspctrm_{band}_{freq}_sub{tag}.mat"))):
                    try:
                        data = _first_mat_variable(f, flatten=True)
                        if data is not None:
                            out[(subj, band, freq, 0)] = data
                    except Exception as e:  # noqa: BLE001
                        logger.warning(f"Error loading {f}: {e}")
    logger.info(f"Loaded {len(out)} EEG power spectrum samples")
    return out


def load_eeg_erp_features(erp_dir, subject_list, band_list, freq_list):
    """:128-186 -- {(subject, band, freq, 0): fp32 array, shape kept} from `ERP_subNN_<band>_<freq>*.mat`: HDF5
    (v7.3) layout first, MATLAB v5 (`scipy.io.loadmat`, first variable) when that fails; a later file matching the
    same key overwrites an earlier one (sorted order)."""
    erp_dir = Path(erp_dir)
    out = {}
    for subj in subject_list:
        tag = f"{subj:02d}"
        for band in band_list:
            for freq in freq_list:
                for f in sorted(glob.glob(str(erp_dir / f"ERP_sub{tag}_{band}_{freq}*.mat"))):
                    try:
                        found, data = _hdf5_erp(f)
                        if found:
                            out[(subj, band, freq, 0)] = data
                    except Exception as e:  # noqa: BLE001 - not HDF5 (or no h5py): MATLAB v5 reader
                        try:
                            data = _first_mat_variable(f, flatten=False)
                            if data is not None:
                                out[(subj, band, freq, 0)] = data
                        except Exception:  # noqa: BLE001
                            logger.warning(f"Error loading ERP {f}: {e}")
    logger.info(f"Loaded {len(out)} EEG ERP samples")
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
