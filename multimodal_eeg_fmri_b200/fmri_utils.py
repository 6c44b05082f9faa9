"""Drop-in for fMRI_CODE/fmri_utils.py: the three models (:23-108) and the per-subject ROI aggregation
arithmetic of load_activation_features (:140-147) as a device op.  CSV parsing is out of scope."""
import torch

from . import functional as XF
from .modules import ActivationEncoder, ConnectivityEncoder, fMRIFusionNet  # noqa: F401

__all__ = ["ActivationEncoder", "ConnectivityEncoder", "fMRIFusionNet", "aggregate_roi_timeseries"]


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
