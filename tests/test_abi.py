"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/xmodal_b200.h declares, and the product path refuses to run without CUDA (no fallback)."""
import ctypes

import pytest
import torch

from multimodal_eeg_fmri_b200 import _lib, ops


def test_library_builds_and_loads():
    from multimodal_eeg_fmri_b200 import build
    path = build.build()
    assert path.exists()
    assert _lib.lib().xm_abi_version() == 1


def test_every_declared_symbol_is_exported():
    decl = _lib.declared_functions()
    assert len(decl) >= 35
    dll = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [name for name in decl if not hasattr(dll, name)]
    assert not missing, f"declared in the header but not exported: {missing}"
    for must in ("xm_bandpower_f32", "xm_conv1d_fwd_f32", "xm_conv1d_dgrad_f32", "xm_conv1d_wgrad_f32",
                 "xm_linear_fwd_f32", "xm_infonce_lse_f32", "xm_infonce_grad_f32", "xm_roi_meanstd_f32"):
        assert must in decl


def test_error_strings():
    dll = _lib.lib()
    assert dll.xm_strerror(0) == b"ok"
    assert b"invalid" in dll.xm_strerror(-1)
    assert b"unsupported" in dll.xm_strerror(-2)


def test_argument_validation_without_gpu():
    # null pointers / bad shapes are rejected before any CUDA call
    dll = _lib.lib()
    assert dll.xm_roi_meanstd_f32(None, 1, 1, 1, None, None) == -1
    assert dll.xm_linear_fwd_f32(None, None, None, None, 4, 4, 4, 4, 4, 4, 0, 0, 1, None, None) == -1
    assert dll.xm_conv1d_wgrad_workspace(8, 64, 64, 7) > 0
    assert dll.xm_bn_nsplit(4096 * 500, 64) >= 1


def test_no_cpu_fallback():
    x = torch.randn(4, 8)
    with pytest.raises(_lib.XmodalError):
        ops.linear_fwd(x, torch.randn(3, 8))
    from multimodal_eeg_fmri_b200.bridge_utils import symmetric_infonce
    with pytest.raises(_lib.XmodalError):
        symmetric_infonce(torch.randn(4, 8), torch.randn(4, 8))
