"""Seeded synthetic inputs of SURVEY.md section 8d (there are no data files: the reference's clinical data is
private).  All generators run on the CPU with an explicit torch.Generator so that the CUDA path,
the oracle and the CPU baseline see identical tensors.

Paired configs (3/4): sample i of EEG and fMRI share a latent z_i ~ N(0, I_16) so that InfoNCE has
signal: EEG = sum_k z_ik * spatial_k(c) * temporal_k(t) + noise; ROI series and connectivity are
linear images of the same z plus noise.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

LATENT = 16


def paired_batch(batch: int, eeg_channels: int = 64, eeg_samples: int = 500, n_roi: int = 200, n_tr: int = 100,
                 conn_dim: int | None = None, seed: int = 42, noise: float = 1.0,
                 offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> eeg (B, C, T), roi_series (B, TR, ROI), conn (B, conn_dim) fp32 CPU tensors.
    `offset` selects which rows of the (conceptually infinite) sample stream this batch holds, so
    rank r of a sharded run draws rows [r*B, (r+1)*B) of the same global batch."""
    conn_dim = n_roi * n_roi if conn_dim is None else conn_dim
    gm = torch.Generator().manual_seed(seed)  # mixing matrices: shared by every batch / rank
    sp = torch.randn(LATENT, eeg_channels, generator=gm) / math.sqrt(LATENT)
    t = torch.arange(eeg_samples, dtype=torch.float32) / max(eeg_samples, 1)
    freqs = torch.arange(1, LATENT + 1, dtype=torch.float32).unsqueeze(1)
    phase = torch.rand(LATENT, 1, generator=gm) * 2 * math.pi
    tb = torch.sin(2 * math.pi * freqs * t.unsqueeze(0) * 3.0 + phase)
    roi_mix = torch.randn(LATENT, n_roi, generator=gm) / math.sqrt(LATENT)
    conn_mix = torch.randn(LATENT, conn_dim, generator=gm) / math.sqrt(LATENT)

    gs = torch.Generator().manual_seed(seed * 1_000_003 + 17 + offset)  # per-batch sample stream
    z = torch.randn(batch, LATENT, generator=gs)
    eeg = torch.einsum("bk,kc,kt->bct", z, sp, tb)
    eeg.add_(torch.randn(batch, eeg_channels, eeg_samples, generator=gs), alpha=noise)
    roi = (z @ roi_mix).unsqueeze(1) + noise * torch.randn(batch, n_tr, n_roi, generator=gs)
    conn = z @ conn_mix
    conn.add_(torch.randn(batch, conn_dim, generator=gs), alpha=noise)
    return eeg.contiguous(), roi.contiguous(), conn.contiguous()


def eeg_recordings(n_rec: int, channels: int = 128, n_samples: int = 8192, fs: float = 1000.0, seed: int = 42):
    """Config 5 recordings: sum of 6/10/20 Hz sinusoids with per-channel amplitude/phase + N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / fs
    x = torch.randn(n_rec, channels, n_samples, generator=g, dtype=torch.float32)
    for f in (6.0, 10.0, 20.0):
        a = torch.rand(n_rec, channels, 1, generator=g) + 0.5
        ph = torch.rand(n_rec, channels, 1, generator=g) * 2 * math.pi
        x += (a.double() * torch.sin(2 * math.pi * f * t + ph.double())).float()
    return x.contiguous()
