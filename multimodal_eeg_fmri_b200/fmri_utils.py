"""Drop-in for fMRI_CODE/fmri_utils.py: the three models (:23-108), the per-subject ROI aggregation
arithmetic of load_activation_features (:140-147) as a device op, and the CSV loaders (:115-241; SURVEY.md
section 8f rank 3): files are parsed on the host, the (TR, ROI) series of ALL subjects are aggregated by one
device launch per distinct shape instead of one NumPy pass per file."""
import logging
from collections import defaultdict
from pathlib import Path

import numpy as np
import pandas as pd
import torch

from . import functional as XF
from .modules import ActivationEncoder, ConnectivityEncoder, fMRIFusionNet  # noqa: F401

logger = logging.getLogger(__name__)

__all__ = ["ActivationEncoder", "ConnectivityEncoder", "fMRIFusionNet", "aggregate_roi_timeseries",
           "load_activation_features", "load_connectivity_features", "load_fmri_labels"]


def aggregate_roi_timeseries(x: torch.Tensor, agg_method: str = "both") -> torch.Tensor:
    """x (B, TR, ROI) CUDA fp32 -> (B, ROI) ['mean' | 'std'] or (B, 2*ROI) ['both'], NaN -> 0 first,
    population std -- the `agg_method` switch of fMRI_CODE/fmri_utils.py:140-149 (same ValueError)."""
    if agg_method not in ("mean", "std", "both"):
        raise ValueError(f"Unknown agg method: {agg_method}")
    if x.dim() != 3:
        raise ValueError("expected (subjects, TR, ROI)")
    roi = x.shape[2]
    if x.shape[0] == 0 or roi == 0:
        both = torch.empty(x.shape[0], 2 * roi, device=x.device, dtype=torch.float32)
    else:
        both = XF.roi_meanstd(x)
    if agg_method == "mean":
        return both[:, :roi]
    if agg_method == "std":
        return both[:, roi:]
    return both


def connectivity_from_timeseries(x: torch.Tensor, prepared: bool = False) -> torch.Tensor:
    """x (B, TR, ROI) CUDA fp32 -> (B, ROI*ROI): the flattened ROI x ROI Pearson correlation matrix of every sample
    (numpy.corrcoef of the columns, NaN -> 0 first) -- fMRIFusionNet's `connectivity` input derived on the device
    from the same series the activation features come from, instead of a 4*ROI*ROI-byte row per sample crossing
    PCIe (the reference reads precomputed matrices from CSV, fmri_utils.py:161-198)."""
    if x.dim() != 3:
        raise ValueError("expected (subjects, TR, ROI)")
    if x.shape[0] == 0 or x.shape[2] == 0:
        return torch.empty(x.shape[0], x.shape[2] * x.shape[2], device=x.device, dtype=torch.float32)
    from . import ops
    return ops.roi_corrcoef(x, prepared)  # prepared: (3B, ROI*ROI) row-stacked tf32 split for the 3-pass projection


# ------------------------------------------------------------------------- CSV loaders (fmri_utils.py:115-241)
def _read_numeric_csv(filepath) -> np.ndarray:
    """One reference CSV as fp32 (rows, columns) with the 'Subject' column dropped and NaN -> 0
    (fmri_utils.py:135-139 / :178-182)."""
    df = pd.read_csv(filepath)
    if "Subject" in df.columns:
        df = df.drop("Subject", axis=1)
    return np.nan_to_num(df.values.astype(np.float32), nan=0.0)


def _device_for(device):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def load_activation_features(data_dir, subject_list, activation_types, agg_method="both", *, device=None):
    """fmri_utils.py:115-158: {subject: fp32 (n_types * [1|2] * ROI,)} from `sub-<s>/subject_<s>_activation_<t>.csv`,
    types concatenated in `activation_types` order, missing files skipped, subjects without any file left out.
    As in the reference an unreadable file or an unknown `agg_method` is logged per file and skipped (the
    ValueError of :149 is raised inside the reference's own try block), so an unknown method yields {}.
    Results are CPU tensors like the reference's; the mean/std pass itself runs on `device` (default: current
    CUDA device) over all files of one (TR, ROI) shape at once."""
    data_dir = Path(data_dir)
    parsed = []  # (subject, position, series)
    for subj in dict.fromkeys(subject_list):  # a repeated subject is loaded once (the reference overwrites its entry)
        subj_dir = data_dir / f"sub-{subj}"
        for pos, act_type in enumerate(activation_types):
            filepath = subj_dir / f"subject_{subj}_activation_{act_type}.csv"
            if not filepath.exists():
                continue
            try:
                data = _read_numeric_csv(filepath)
                if agg_method not in ("mean", "std", "both"):
                    raise ValueError(f"Unknown agg method: {agg_method}")
                if data.ndim != 2 or data.shape[0] == 0:
                    raise ValueError(f"no rows in {filepath}")
                parsed.append((subj, pos, data))
            except Exception as e:  # noqa: BLE001 - reference behaviour: log and continue
                logger.warning(f"Error loading {filepath}: {e}")
    by_shape = defaultdict(list)
    for i, (_, _, data) in enumerate(parsed):
        by_shape[data.shape].append(i)
    dev = _device_for(device)
    agg = [None] * len(parsed)
    for shape, idx in by_shape.items():
        host = torch.from_numpy(np.stack([parsed[i][2] for i in idx]))
        if dev.type == "cuda":
            host = host.pin_memory()
        out = aggregate_roi_timeseries(host.to(dev, non_blocking=True), agg_method).cpu()
        for row, i in enumerate(idx):
            agg[i] = out[row]
    per_subject = defaultdict(list)
    for (subj, _, _), a in zip(parsed, agg):  # already in subject, then activation_types order
        per_subject[subj].append(a)
    features = {subj: torch.cat(per_subject[subj]).to(torch.float32).clone() for subj in subject_list if subj in per_subject}
    logger.info(f"fMRI activation features: {len(features)}/{len(subject_list)} subjects")
    if features:
        logger.info(f"  Activation feature dim: {next(iter(features.values())).shape[0]}")
    return features


def load_connectivity_features(data_dir, subject_list, connectivity_types):
    """fmri_utils.py:161-198: {subject: fp32 flattened connectivity matrices, types concatenated} from
    `sub-<s>/subject_<s>_fdr_PPI_Connectivity_<t>.csv` (pure byte movement: host only, bit-exact)."""
    data_dir = Path(data_dir)
    features = {}
    for subj in subject_list:
        subj_features = []
        subj_dir = data_dir / f"sub-{subj}"
        for conn_type in connectivity_types:
            filepath = subj_dir / f"subject_{subj}_fdr_PPI_Connectivity_{conn_type}.csv"
            if not filepath.exists():
                continue
            try:
                subj_features.append(_read_numeric_csv(filepath).flatten())
            except Exception as e:  # noqa: BLE001
                logger.warning(f"Error loading {filepath}: {e}")
        if subj_features:
            features[subj] = torch.tensor(np.concatenate(subj_features), dtype=torch.float32)
    logger.info(f"fMRI connectivity features: {len(features)}/{len(subject_list)} subjects")
    if features:
        logger.info(f"  Connectivity feature dim: {next(iter(features.values())).shape[0]}")
    return features


_SUBJECT_COLUMNS = ["Subject", "subject", "SubjectID", "ID", "id"]
_LABEL_COLUMNS = ["Label", "label", "Outcome", "outcome", "Class", "class", "Group", "group"]


def load_fmri_labels(label_path, subject_list):
    """fmri_utils.py:201-241: first existing of labels.csv / outcomes.csv / subjects_labels.csv / ../labels.csv;
    string labels 'good' | 'positive' | 'yes' | '1' (any case) -> 1 else 0; numeric labels int(); subjects outside
    `subject_list` dropped; ValueError when the columns cannot be identified; random dummy labels when no file."""
    label_path = Path(label_path)
    candidates = [label_path / "labels.csv", label_path / "outcomes.csv", label_path / "subjects_labels.csv",
                  label_path.parent / "labels.csv"]
    label_file = next((lf for lf in candidates if lf.exists()), None)
    if label_file is None:
        logger.warning("No fMRI label file found. Using dummy labels.")
        return {subj: np.random.randint(0, 2) for subj in subject_list}
    df = pd.read_csv(label_file)
    subj_col = next((c for c in _SUBJECT_COLUMNS if c in df.columns), None)
    label_col = next((c for c in _LABEL_COLUMNS if c in df.columns), None)
    if not subj_col or not label_col:
        raise ValueError(f"Cannot identify columns in {label_file}: {df.columns.tolist()}")
    class_labels = {}
    for subj, label in zip(df[subj_col].tolist(), df[label_col].tolist()):
        subj = int(subj)
        if subj not in subject_list:
            continue
        if isinstance(label, str):
            label = 1 if label.lower() in ["good", "positive", "yes", "1"] else 0
        else:
            label = int(label)
        class_labels[subj] = label
    logger.info(f"fMRI labels: {len(class_labels)} subjects, classes={set(class_labels.values())}")
    return class_labels
