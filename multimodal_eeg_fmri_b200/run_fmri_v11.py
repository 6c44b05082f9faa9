"""Drop-in for the hot-path part of fMRI_CODE/run_fmri_v11.py (config 2): `train_epoch` (:430-450),
`evaluate` (:453-...) and a `main()` on the synthetic tensors of SURVEY.md section 8d (200 ROI x 100 TR, batch 64).
(The reference file does not parse -- SyntaxError at :44, SURVEY.md section 0 -- and its models are the ones in
fmri_utils.py.)"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .fmri_utils import ActivationEncoder, ConnectivityEncoder, aggregate_roi_timeseries, fMRIFusionNet  # noqa: F401

__all__ = ["train_epoch", "evaluate", "run_experiment", "main", "fMRIFusionNet"]


def train_epoch(model, train_loader, optimizer, criterion, device, task="classification", grad_clip=1.0):
    model.train()
    total = 0.0
    for activation, connectivity, class_labels, reg_labels, _ in train_loader:
        activation, connectivity = activation.to(device), connectivity.to(device)
        labels = (class_labels if task == "classification" else reg_labels).to(device)
        optimizer.zero_grad()
        loss = criterion(model(activation, connectivity), labels)
        loss.backward()
        if grad_clip > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
        optimizer.step()
        total += loss.item()
    return total / len(train_loader)


@torch.no_grad()
def evaluate(model, data_loader, device, task="classification", num_classes=2):
    """Predictions / targets / probabilities of one pass in eval mode."""
    model.eval()
    preds, targets, probs = [], [], []
    for activation, connectivity, class_labels, reg_labels, _ in data_loader:
        out = model(activation.to(device), connectivity.to(device))
        if task == "classification":
            probs.append(F.softmax(out, dim=1).cpu())
            preds.append(out.argmax(dim=1).cpu())
            targets.append(class_labels)
        else:
            preds.append(out.cpu())
            targets.append(reg_labels)
    cat = lambda xs: torch.cat(xs) if xs else torch.empty(0)
    return {"preds": cat(preds), "targets": cat(targets), "probs": cat(probs)}


def run_experiment(dataset, config=None, task="classification", epochs: int = 1, device: str = "cuda"):
    """One training run over `dataset` (a list of pre-collated batches) with the reference's
    hyper-parameters (AdamW lr 1e-4, wd 1e-4, clip 1.0); the CV / plotting shell is out of scope."""
    cfg = dict(lr=1e-4, weight_decay=1e-4, grad_clip=1.0, hidden_dim=64, dropout=0.4)
    cfg.update(config or {})
    act0, conn0 = dataset[0][0], dataset[0][1]
    model = fMRIFusionNet(act0.shape[1], conn0.shape[1], cfg["hidden_dim"], 2, cfg["dropout"], task).to(device)
    opt = torch.optim.AdamW(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    crit = torch.nn.CrossEntropyLoss() if task == "classification" else torch.nn.MSELoss()
    losses = [train_epoch(model, dataset, opt, crit, device, task, cfg["grad_clip"]) for _ in range(epochs)]
    return model, losses


def main(steps: int = 10, batch: int = 64, n_roi: int = 200, n_tr: int = 100, seed: int = 42, device: str = "cuda"):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    roi = torch.randn(batch, n_tr, n_roi, generator=g).to(device)
    act = aggregate_roi_timeseries(roi, "both").cpu()
    conn = torch.randn(batch, n_roi * n_roi, generator=g)
    y = torch.randint(0, 2, (batch,), generator=g)
    dataset = [(act, conn, y, y.float(), list(range(batch)))]
    model, losses = run_experiment(dataset, None, "classification", epochs=steps, device=device)
    print(f"run_fmri_v11 (synthetic, batch {batch}): loss {losses[0]:.4f} -> {losses[-1]:.4f}")
    return losses


if __name__ == "__main__":
    main()
