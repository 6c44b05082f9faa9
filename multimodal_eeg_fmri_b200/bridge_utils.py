"""Drop-in for bridge_utils.py: the bridge model (:22-114), the aligned feature dataset (:120-152), the
attribution helpers that follow training in every reference pipeline (:158-270) and
the contrastive functions the north-star adds to this module (`similarity_matrix`, `symmetric_infonce`; no
reference implementation -- SURVEY.md section 0)."""
import logging
from typing import List

import numpy as np
import torch
from torch.utils.data import Dataset

from . import functional as XF
from .modules import EEGfMRIBridgeFusionNet, LearnedFusionModule  # noqa: F401

logger = logging.getLogger(__name__)

__all__ = ["EEGfMRIBridgeFusionNet", "BridgeFeatureDataset", "BridgeRawDataset", "collate_bridge", "similarity_matrix", "symmetric_infonce",
           "BridgeGradientSaliency", "BridgeIntegratedGradients", "extract_attention_and_fusion_weights"]

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


class BridgeRawDataset(Dataset):
    """_test_bridge.py:391-453: subject alignment of the RAW feature dicts.  Every ERP entry (subject, band, freq,
    label) is paired with the power entry of the same key and the connectivity entry (subject, band.lower(), cond,
    label) of the first condition in `func_segments` that exists; a missing power / connectivity entry is replaced
    by zeros shaped like the FIRST entry of the respective dict (dropped if that dict is empty).  A subject is kept
    when it has EEG samples, both fMRI dicts and a label; item = (eeg_samples, fmri_act, fmri_conn, label, subject).
    (`bands` is accepted and unused, as in the reference.)"""

    def __init__(self, eeg_erp, eeg_pw, eeg_conn, fmri_act, fmri_conn, labels, subject_list, bands, func_segments):
        # shapes of the zero stand-ins; a FRESH array per missing entry as in the reference (:416-421), so samples
        # never alias each other (an in-place normalisation of one padded sample must not touch the others)
        pw_shape = next((v.shape for v in eeg_pw.values()), None)
        conn_shape = next((v.shape for v in eeg_conn.values()), None)
        eeg_by_subject = {}
        for key, erp in eeg_erp.items():
            sid = key[0] if isinstance(key[0], int) else int(key[0])
            pw = eeg_pw.get(key)
            if pw is None and pw_shape is not None:
                pw = np.zeros(pw_shape, dtype=np.float32)
            band = str(key[1]).lower()
            hit = next((c for c in func_segments if (key[0], band, c, key[3]) in eeg_conn), None)
            if hit is not None:
                conn = eeg_conn[(key[0], band, hit, key[3])]
            else:
                conn = None if conn_shape is None else np.zeros(conn_shape, dtype=np.float32)
            if pw is not None and conn is not None:
                eeg_by_subject.setdefault(sid, []).append((erp, pw, conn))
        self.samples = []
        for subj in sorted(subject_list):
            sid = int(subj)
            have = {"EEG": sid in eeg_by_subject, "fMRI-Act": sid in fmri_act, "fMRI-Conn": sid in fmri_conn, "Label": sid in labels}
            if all(have.values()):
                self.samples.append({"subject": sid, "label": labels[sid], "eeg_samples": eeg_by_subject[sid],
                                     "fmri_act": fmri_act[sid], "fmri_conn": fmri_conn[sid]})
            else:
                logger.debug("Subject %d excluded. Missing: %s", sid, ", ".join(k for k, v in have.items() if not v))
        if not self.samples:
            logger.error("!!! NO ALIGNED SUBJECTS FOUND !!! Check Subject IDs and file paths.")
            return
        counts = [len(s["eeg_samples"]) for s in self.samples]
        logger.info("BridgeRawDataset: %d aligned subjects (EEG samples per subject: min=%d, max=%d)", len(self.samples),
                    min(counts), max(counts))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        s = self.samples[idx]
        return s["eeg_samples"], s["fmri_act"], s["fmri_conn"], s["label"], s["subject"]


def collate_bridge(batch):
    """_test_bridge.py:755-760."""
    eeg = torch.stack([b[0] for b in batch])
    fmri = torch.stack([b[1] for b in batch])
    labels = torch.tensor([b[2] for b in batch], dtype=torch.long)
    return eeg, fmri, labels, [b[3] for b in batch]


# ------------------------------------------------------------------------- attribution (bridge_utils.py:158-270)
def _class_logit_gradients(model, eeg, fmri, target_class, target_rows=None):
    """d(logit[target])/d(eeg, fmri) for every row of ONE batch (one forward, one backward through the fused
    kernels).  target_class None -> argmax of the logits of the first `target_rows` rows, tiled over the batch."""
    eeg = eeg.clone().detach().requires_grad_(True)
    fmri = fmri.clone().detach().requires_grad_(True)
    logits = model(eeg, fmri)
    if target_class is None:
        head = logits if target_rows is None else logits[:target_rows]
        target_class = head.argmax(dim=1)
    target_class = target_class.to(logits.device).view(-1)
    if target_class.numel() != logits.shape[0]:
        target_class = target_class.repeat(logits.shape[0] // target_class.numel())
    model.zero_grad()
    one_hot = torch.zeros_like(logits)
    one_hot.scatter_(1, target_class.view(-1, 1), 1)
    logits.backward(gradient=one_hot)
    return eeg.grad, fmri.grad, target_class


class BridgeGradientSaliency:
    """bridge_utils.py:158-183: |d logit[target] / d input| per sample (target = predicted class by default)."""

    def __init__(self, model, device):
        self.model = model
        self.device = device

    def compute(self, eeg_feats, fmri_feats, target_class=None):
        self.model.eval()
        ge, gf, _ = _class_logit_gradients(self.model, eeg_feats.to(self.device), fmri_feats.to(self.device), target_class)
        return {"eeg": ge.abs().cpu().numpy(), "fmri": gf.abs().cpu().numpy()}


class BridgeIntegratedGradients:
    """bridge_utils.py:190-229 with the `n_steps` interpolation points of all samples run as ONE batch of
    n_steps * B rows (the reference runs n_steps sequential forward/backward passes): rows [a*B, (a+1)*B) hold
    alpha_a * x.  As in the reference, an unspecified target class is fixed by the FIRST interpolation point
    (alpha = 0, the all-zero baseline -- bridge_utils.py:213-214 assigns `target_class` inside the loop once) and
    the attribution is |x * mean_alpha grad|."""

    def __init__(self, model, device, n_steps=50):
        self.model = model
        self.device = device
        self.n_steps = n_steps

    def compute(self, eeg_feats, fmri_feats, target_class=None):
        self.model.eval()
        eeg = eeg_feats.to(self.device)
        fmri = fmri_feats.to(self.device)
        B, n = eeg.shape[0], self.n_steps
        # np.linspace alphas enter the fp32 product as scalars (`baseline + alpha * diff`): rounded to fp32 first
        alphas = torch.tensor(np.linspace(0, 1, n), dtype=eeg.dtype, device=eeg.device)
        scale = lambda x: (alphas.view(n, 1, 1) * x.unsqueeze(0)).reshape(n * B, -1)  # noqa: E731
        ge, gf, _ = _class_logit_gradients(self.model, scale(eeg), scale(fmri), target_class, target_rows=B)
        eeg_ig = eeg * ge.view(n, B, -1).mean(0)
        fmri_ig = fmri * gf.view(n, B, -1).mean(0)
        return {"eeg": eeg_ig.abs().cpu().numpy(), "fmri": fmri_ig.abs().cpu().numpy()}


def extract_attention_and_fusion_weights(model, dataset, device):
    """bridge_utils.py:236-270: per subject {subject, label, prediction, fusion_weights (2,), attn_weights (2,)};
    the whole dataset goes through the model as one batch (eval mode: rows are independent)."""
    model.eval()
    if len(dataset) == 0:
        return []
    rows = [dataset[i] for i in range(len(dataset))]
    with torch.no_grad():
        eeg = torch.stack([r[0] for r in rows]).to(device)
        fmri = torch.stack([r[1] for r in rows]).to(device)
        logits, _, fusion_w, attn_w = model(eeg, fmri, return_features=True, return_weights=True)
        pred = logits.argmax(dim=1).cpu().numpy()
        fusion_w = fusion_w.cpu().numpy()
        attn_w = attn_w.cpu().numpy()
    return [{"subject": r[3], "label": r[2], "prediction": int(pred[i]), "fusion_weights": fusion_w[i].squeeze(),
             "attn_weights": attn_w[i].squeeze()} for i, r in enumerate(rows)]
