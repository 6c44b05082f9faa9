"""xmodal-b200: the paired EEG/fMRI cross-modal training step of bacon205/Multimodal_eeg_fmri as
hand-written sm_100a kernels behind the reference's own Python API.

    from multimodal_eeg_fmri_b200.bridge_utils import EEGfMRIBridgeFusionNet, symmetric_infonce
    from multimodal_eeg_fmri_b200.enhanced_models_v4 import EnhancedERPEncoder
    from multimodal_eeg_fmri_b200.fmri_utils import fMRIFusionNet
    from multimodal_eeg_fmri_b200.eeg_data_utils import band_power, window_indices

Importing the package does not load the CUDA library; the first op does, and raises if it is missing
(`python -m multimodal_eeg_fmri_b200.build`).  There is no CPU fallback.
"""
__version__ = "0.1.0"
