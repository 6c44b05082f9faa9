"""Drop-in for the hot-path part of EEG_CODE/run_training_lite.py (config 1): the model wrapper
(:302-328, argument order forward(pw, erp, conn)), `collate_balanced` (:331-350), the per-batch step
of the training loop (:478-489; AdamW lr 5e-5 / wd 0.01, label smoothing 0.1, clip 1.0) and a
`main()` that runs it.  The reference's main() reads private .mat files through a Config that does
not match the script; here main() trains on the seeded synthetic tensors of
SURVEY.md section 8d (64 ch x 500 samples, conn 6048, batch 32)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .eeg_data_utils import normalize_modality  # noqa: F401  (re-exported: run_training_lite.py:48-51)
from .modules import EnhancedTriModalFusionNetV4Lite, LabelSmoothingCrossEntropy

__all__ = ["ImprovedTriModalFusionNetLite", "collate_balanced", "get_lite_fusion_weights", "train_one_epoch", "main",
           "normalize_modality"]


def get_lite_fusion_weights(model):
    """crossmodal_v4_enhancements.py:1146-1152."""
    if hasattr(model, "get_fusion_weights"):
        return model.get_fusion_weights()
    return getattr(model, "_fusion_weights", None)


class ImprovedTriModalFusionNetLite(nn.Module):
    """run_training_lite.py:302-328.  NOTE the argument order: forward(pw, erp, conn)."""

    def __init__(self, in_pw_dim, in_erp_dim, in_conn_dim, fusion_dim=96, num_classes=2, dropout=0.4, conn_boost=1.3):
        super().__init__()
        self.model = EnhancedTriModalFusionNetV4Lite(erp_channels=in_erp_dim, pw_channels=in_pw_dim,
                                                     conn_features=in_conn_dim, hidden_dim=fusion_dim,
                                                     num_classes=num_classes, dropout=dropout, conn_boost=conn_boost)
        self.fusion_weight_history = []

    def forward(self, pw, erp, conn):
        logits, _ = self.model(erp, pw, conn, return_fusion_weights=True)
        return logits

    def get_fusion_weights(self):
        return get_lite_fusion_weights(self.model)

    def track_fusion_weights(self):
        w = self.get_fusion_weights()
        if w:
            self.fusion_weight_history.append(dict(w))  # a snapshot: the live object reads its device tensor on access


def collate_balanced(batch):
    """run_training_lite.py:331-350: dict or tuple samples -> (erp, pw, conn, labels[int64], subjects)."""
    cols = ([], [], [], [], [])
    for s in batch:
        vals = (s["erp"], s["pw"], s["conn"], s["label"], s["subject"]) if isinstance(s, dict) else s[:5]
        for c, v in zip(cols, vals):
            c.append(v)
    erp, pw, conn, labels, subj = cols
    return torch.stack(erp), torch.stack(pw), torch.stack(conn), torch.tensor(labels, dtype=torch.long), subj


def train_one_epoch(model, train_loader, optimizer, criterion, device, grad_clip: float = 1.0) -> float:
    """The body of run_training_lite.py:474-489 for one epoch; returns the summed loss."""
    model.train()
    total = 0.0
    for erp, pw, conn, y, _ in train_loader:
        erp, pw, conn, y = erp.to(device), pw.to(device), conn.to(device), y.to(device)
        optimizer.zero_grad()
        loss = criterion(model(pw, erp, conn), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=grad_clip)
        optimizer.step()
        total += loss.item()
    return total


def main(steps: int = 10, batch: int = 32, channels: int = 64, samples: int = 500, conn_dim: int = 6048,
         lr: float = 5e-5, seed: int = 42, device: str = "cuda"):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    model = ImprovedTriModalFusionNetLite(channels, channels, conn_dim).to(device)
    crit = LabelSmoothingCrossEntropy(smoothing=0.1)
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=0.01)
    data = [(torch.randn(channels, samples, generator=g), torch.randn(channels, samples, generator=g),
             torch.randn(conn_dim, generator=g), int(torch.randint(0, 2, (1,), generator=g)), i) for i in range(batch)]
    loader = [collate_balanced(data)]
    losses = [train_one_epoch(model, loader, opt, crit, device) for _ in range(steps)]
    print(f"run_training_lite (synthetic, batch {batch}): loss {losses[0]:.4f} -> {losses[-1]:.4f}")
    return losses


if __name__ == "__main__":
    main()
