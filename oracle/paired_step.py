"""CPU oracle (test infrastructure): the composite paired EEG/fMRI training step of SURVEY.md section 3 (E).

The reference has no single call site for this step: its pieces are the reference
modules restated in `oracle.models` (PINNED against tests/golden/*.npz) joined by the authored
InfoNCE definition of `oracle.infonce` (PARITY UNPINNED).  The step recipe -- zero_grad, backward,
clip_grad_norm_(1.0), AdamW(lr 1e-4, wd 1e-4) -- is the one of _test_bridge.py:775-788,869.

The state dict uses the key names of multimodal_eeg_fmri_b200.training.PairedBridgeModel:
`eeg_encoder.*` (EnhancedERPEncoder | LiteERPEncoder), `fmri_net.*` (fMRIFusionNet), `bridge.*`
(EEGfMRIBridgeFusionNet) -- each sub-dict is exactly the reference class's state_dict.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import infonce as oi
from . import models as om

SD = Dict[str, torch.Tensor]


def paired_embeddings(P: SD, eeg, roi_series, conn, encoder: str = "v4", nhead: int = 4, train: bool = True):
    """eeg (B, C, T), roi_series (B, TR, ROI), conn (B, conn_dim) | None -> the two (B, bridge_dim) embeddings.

    EEG encoder (enhanced_models_v4.py:114-193 or crossmodal_v4_enhancements.py:817-845) ->
    eeg_proj (bridge_utils.py:34-39,71); ROI mean/std (fmri_utils.py:140-147) -> fMRIFusionNet fused
    feature (fmri_utils.py:90-100, return_features) -> fmri_proj (bridge_utils.py:40-45,72)."""
    if encoder == "v4":
        eeg_feat = om.enhanced_erp_encoder(P, "eeg_encoder.", eeg, nhead=nhead, train=train)
    elif encoder == "lite":
        eeg_feat = om.lite_encoder(P, "eeg_encoder.", eeg, train=train)
    else:
        raise ValueError(encoder)
    act = om.roi_meanstd(roi_series)
    if conn is None:  # connectivity derived from the ROI series
        conn = om.roi_connectivity(roi_series)
    _, fmri_feat = om.fmri_fusion_net(P, "fmri_net.", act, conn, train=train)
    return om.bridge_projections(P, "bridge.", eeg_feat, fmri_feat)


def paired_loss(P: SD, eeg, roi_series, conn, temperature: float = 0.07, encoder: str = "v4", train: bool = True):
    e, f = paired_embeddings(P, eeg, roi_series, conn, encoder, train=train)
    return oi.symmetric_infonce(e, f, temperature)


def trainable_keys(P: SD, reached_only: bool = True) -> List[str]:
    """Floating-point entries that are parameters (not BN running stats / PE table); with
    `reached_only`, only those the InfoNCE loss reaches (the supervised heads get no gradient)."""
    skip = ("running_mean", "running_var", "num_batches_tracked", "pos_encoder.pe")
    keys = [k for k, v in P.items() if v.is_floating_point() and not k.endswith(skip)]
    if reached_only:
        dead = ("fmri_net.head.", "bridge.cross_attn.", "bridge.fusion.", "bridge.classifier.")
        keys = [k for k in keys if not k.startswith(dead)]
    return keys


def bias_before_batchnorm_keys(P: SD) -> List[str]:
    """Biases of a Conv1d / Linear that feeds a train-mode BatchNorm directly: their gradient is
    mathematically zero (BN subtracts the batch mean), so Adam turns pure rounding noise into +-lr
    steps of arbitrary sign.  Parameter comparisons after optimizer steps skip these."""
    out = []
    for k in trainable_keys(P):
        if not k.endswith(".bias"):
            continue
        head, idx = k[: -len(".bias")].rsplit(".", 1)
        if idx.isdigit() and f"{head}.{int(idx) + 1}.running_mean" in P:
            out.append(k)
    return out


def paired_loss_and_grads(P: SD, eeg, roi_series, conn, temperature: float = 0.07, encoder: str = "v4"):
    keys = trainable_keys(P)
    leaves = {k: P[k].detach().clone().requires_grad_(True) for k in keys}
    loss = paired_loss({**P, **leaves}, eeg, roi_series, conn, temperature, encoder, train=True)
    grads = dict(zip(keys, torch.autograd.grad(loss, [leaves[k] for k in keys])))
    return loss.detach(), grads


def paired_train_step(P: SD, state: dict, eeg, roi_series, conn, temperature: float = 0.07, encoder: str = "v4",
                      lr: float = 1e-4, weight_decay: float = 1e-4, max_norm: float = 1.0):
    """One full step; mutates P / state in place (parameters only -- BN running statistics are not
    part of the compared quantities here).  Returns (loss, pre-clip grad norm)."""
    loss, grads = paired_loss_and_grads(P, eeg, roi_series, conn, temperature, encoder)
    keys = list(grads)
    with torch.no_grad():
        total = om.clip_and_adamw({k: P[k] for k in keys}, grads, state, lr, weight_decay, max_norm)
    return loss, total


def sharded_paired_loss(P: SD, shards: Sequence[tuple], temperature: float = 0.07, encoder: str = "v4"):
    """Definition of the data-parallel step's loss: the global-batch loss over the concatenated
    shards (SyncBN statistics + global negatives).  `shards` = [(eeg_r, roi_r, conn_r), ...]."""
    eeg = torch.cat([s[0] for s in shards], 0)
    roi = torch.cat([s[1] for s in shards], 0)
    conn = None if shards[0][2] is None else torch.cat([s[2] for s in shards], 0)
    return paired_loss(P, eeg, roi, conn, temperature, encoder)
