"""Shared pytest configuration: the `gpu` marker, golden-fixture loading, comparison helpers."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> dict with 'sd' (OrderedDict of tensors), 'inputs', 'outputs', 'grads', 'in_grads', raw npz."""
    z = np.load(GOLDEN / f"{name}.npz")
    out = {"sd": {}, "sd_after": {}, "grads": {}, "inputs": {}, "outputs": {}, "in_grads": {}, "raw": z}
    for k in z.files:
        head, _, rest = k.partition("/")
        t = torch.from_numpy(z[k])
        if head == "sd":
            out["sd"][rest] = t
        elif head == "sd_after":
            out["sd_after"][rest] = t
        elif head == "grad":
            out["grads"][rest] = t
        elif head == "in":
            out["inputs"][int(rest)] = t
        elif head == "out":
            out["outputs"][int(rest)] = t
        elif head == "in_grad":
            out["in_grads"][int(rest)] = t
    return out


def _t(a):
    return torch.as_tensor(a).detach().double().cpu()


def rel_err(a, b):
    a, b = _t(a), _t(b)
    return float((a - b).norm() / (b.norm() + 1e-30))


def assert_close_rel(a, b, tol, what="", atol=0.0):
    """||a - b|| <= tol * ||b|| + atol * sqrt(numel).  `atol` is the absolute noise floor for
    quantities that are mathematically zero (e.g. the gradient of a conv bias in front of a
    train-mode BatchNorm), where a relative error is meaningless."""
    a, b = _t(a), _t(b)
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} != {tuple(b.shape)}"
    err = float((a - b).norm())
    bound = tol * float(b.norm()) + atol * (b.numel() ** 0.5)
    if os.environ.get("XM_PRINT_ERRS"):  # tolerance audit: every comparison's measured error next to its bound
        print(f"[err] {what}: rel {err / (float(b.norm()) + 1e-30):.3e} (tol {tol:.1e}, used {err / (bound + 1e-300):.2f} of the bound)")
    assert err <= bound, f"{what}: ||a-b|| = {err:.3e} > {bound:.3e} (rel {err / (float(b.norm()) + 1e-30):.3e}, tol {tol:.1e})"


def bias_before_batchnorm(keys):
    """Keys `<head>.<i>.bias` of a Conv1d/Linear directly followed by a BatchNorm (`<head>.<i+1>.running_mean`
    exists): their gradient is mathematically zero in train mode (BN subtracts the batch mean), so only an
    absolute noise bound is meaningful for them."""
    keys = set(keys)
    out = set()
    for k in keys:
        if not k.endswith(".bias"):
            continue
        head, _, idx = k[: -len(".bias")].rpartition(".")
        if idx.isdigit() and f"{head}.{int(idx) + 1}.running_mean" in keys:
            out.add(k)
    return out


def assert_zero_grad_noise(g_bias, g_weight, what="", frac=2e-3):
    """|g_bias| must be rounding noise: far below the scale of the layer's weight gradient."""
    gb, gw = _t(g_bias), _t(g_weight)
    scale = float(gw.abs().sum()) / max(gb.numel(), 1)  # mean absolute row sum of the weight gradient
    assert float(gb.abs().max()) <= frac * scale + 1e-7, f"{what}: |grad| {float(gb.abs().max()):.3e} vs scale {scale:.3e}"
