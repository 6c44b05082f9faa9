"""Drop-in for EEG_CODE/enhanced_models_v4.py (hot-path classes only; the GNN / Optuna parts are out of
scope)."""
from .modules import (EnhancedERPEncoder, EnhancedPowerEncoder, LearnedFusionModule, PositionalEncoding,  # noqa: F401
                      TemporalTransformerBlock)

__all__ = ["PositionalEncoding", "TemporalTransformerBlock", "EnhancedERPEncoder", "EnhancedPowerEncoder",
           "LearnedFusionModule"]
