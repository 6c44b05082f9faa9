"""Drop-in for EEG_CODE/crossmodal_v4_enhancements.py (hot-path classes)."""
from .modules import (EnhancedConnEncoder, EnhancedERPEncoder, EnhancedPowerEncoder,  # noqa: F401
                      EnhancedTriModalFusionNetV4Lite, HybridFusionModule, LabelSmoothingCrossEntropy,
                      LearnedFusionModule, LiteERPEncoder, LitePowerEncoder, PositionalEncoding,
                      TemporalTransformerBlock)

__all__ = ["PositionalEncoding", "TemporalTransformerBlock", "EnhancedERPEncoder", "EnhancedPowerEncoder",
           "LearnedFusionModule", "LabelSmoothingCrossEntropy", "EnhancedConnEncoder", "HybridFusionModule",
           "LiteERPEncoder", "LitePowerEncoder", "EnhancedTriModalFusionNetV4Lite"]
