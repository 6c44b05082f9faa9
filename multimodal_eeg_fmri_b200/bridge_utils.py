"""Drop-in for bridge_utils.py: the bridge model (:22-114), the aligned feature dataset (:120-152) and
the contrastive functions the north-star adds to this module (`similarity_matrix`,
`symmetric_infonce`; no reference implementation -- SURVEY.md section 0)."""
import logging
from typing import List

import torch
from torch.utils.data import Dataset

from . import functional as XF
from .modules import EEGfMRIBridgeFusionNet, LearnedFusionModule  # noqa: F401

logger = logging.getLogger(__name__)

__all__ = ["EEGfMRIBridgeFusionNet", "BridgeFeatureDataset", "collate_bridge", "similarity_matrix", "symmetric_infonce"]

DEFAULT_TEMPERATURE = 0.07


def similarity_matrix(e: torch.Tensor, f: torch.Tensor, temperature: float = DEFAULT_TEMPERATURE) -> torch.Tensor:
    """S = normalize(e) @ normalize(f)^T / temperature on the tcgen05 GEMM, materialised (Be, Bf)."""
    return XF.similarity_matrix(e, f, temperature)


def symmetric_infonce(e: torch.Tensor, f: torch.Tensor, temperature: float = DEFAULT_TEMPERATURE) -> torch.Tensor:
    """Symmetric InfoNCE over paired rows of e and f (global negatives under data parallelism).
    Differentiable; returns this rank's share of the global-batch loss (the whole loss on 1 GPU)."""
    return XF.symmetric_infonce(e, f, temperature)


class BridgeFeatureDataset(Dataset):
    """bridge_utils.py:120-152: subjects present in all three dicts, int-keyed, iterated in sorted order."""

    def __init__(self, eeg_features, fmri_features, labels, subject_list: List):
        eeg = {int(k): v for k, v in eeg_features.items()}
        fmri = {int(k): v for k, v in fmri_features.items()}
        lab = {int(k): v for k, v in labels.items()}
        self.samples = []
        for subj in sorted(subject_list):
            sid = int(subj)
            if sid in eeg and sid in fmri and sid in lab:
                self.samples.append({"eeg": eeg[sid], "fmri": fmri[sid], "label": lab[sid], "subject": sid})
        if not self.samples:
            logger.error("!!! NO SAMPLES ALIGNED !!! Check subject IDs in EEG and fMRI feature dicts.")
        else:
            logger.info("BridgeFeatureDataset: %d aligned samples found.", len(self.samples))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        s = self.samples[idx]
        return s["eeg"], s["fmri"], s["label"], s["subject"]


def collate_bridge(batch):
    """_test_bridge.py:755-760."""
    eeg = torch.stack([b[0] for b in batch])
    fmri = torch.stack([b[1] for b in batch])
    labels = torch.tensor([b[2] for b in batch], dtype=torch.long)
    return eeg, fmri, labels, [b[3] for b in batch]
