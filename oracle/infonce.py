"""Authored oracle (test infrastructure): EEG<->fMRI similarity matrix + symmetric InfoNCE.

PARITY UNPINNED: bridge_utils.py has no contrastive loss (its model is trained with a class-
weighted nn.CrossEntropyLoss, _test_bridge.py:856-858).  The definition below is the CLIP-style
symmetric cross-entropy of SURVEY.md section 8a row 16, attached to the two shared-space projections
`eeg_proj` / `fmri_proj` (bridge_utils.py:34-45,71-72).  Pinned by tests/test_oracle_infonce.py
(orthonormal rows, all-equal rows, fp64 gradcheck, closed-form gradient).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DEFAULT_TEMPERATURE = 0.07


def similarity_matrix(e: torch.Tensor, f: torch.Tensor, temperature: float = DEFAULT_TEMPERATURE) -> torch.Tensor:
    """S = normalize(e) @ normalize(f)^T / temperature, shape (Be, Bf)."""
    en = F.normalize(e, dim=1)
    fn = F.normalize(f, dim=1)
    return en @ fn.t() / temperature


def symmetric_infonce(e: torch.Tensor, f: torch.Tensor, temperature: float = DEFAULT_TEMPERATURE) -> torch.Tensor:
    """L = 0.5 * [CE(S, arange) + CE(S^T, arange)], mean over the batch; sample i of e pairs with i of f."""
    S = similarity_matrix(e, f, temperature)
    target = torch.arange(S.shape[0], device=S.device)
    return 0.5 * (F.cross_entropy(S, target) + F.cross_entropy(S.t(), target))


def infonce_grad_S(S: torch.Tensor) -> torch.Tensor:
    """Closed form dL/dS = (softmax_row(S) + softmax_col(S) - 2I) / (2B)."""
    B = S.shape[0]
    return (torch.softmax(S, dim=1) + torch.softmax(S, dim=0) - 2 * torch.eye(B, dtype=S.dtype, device=S.device)) / (2 * B)


def sharded_symmetric_infonce(e_shards, f_shards, temperature: float = DEFAULT_TEMPERATURE):
    """Reference semantics of the data-parallel loss: concatenating the per-rank shards and
    evaluating the global loss.  Returns (global loss, list of per-rank loss contributions that
    sum to it) -- the quantity every rank's all-reduced loss must equal."""
    e = torch.cat(list(e_shards), 0)
    f = torch.cat(list(f_shards), 0)
    S = similarity_matrix(e, f, temperature)
    B = S.shape[0]
    lse_r = torch.logsumexp(S, dim=1)
    lse_c = torch.logsumexp(S, dim=0)
    d = torch.diagonal(S)
    per_row = 0.5 * ((lse_r - d) + (lse_c - d)) / B
    parts, o = [], 0
    for es in e_shards:
        parts.append(per_row[o:o + es.shape[0]].sum())
        o += es.shape[0]
    return per_row.sum(), parts
