"""The oracle restatement (oracle/models.py) against golden vectors produced by the REAL reference
classes (oracle/make_golden.py).  CPU only.  Tolerance 1e-5 relative (same fp32 arithmetic, different
op grouping)."""
import json

import pytest
import torch

from conftest import GOLDEN, assert_close_rel, load_golden
from oracle import models as om

TOL = 1e-5


def _params(g):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in g["sd"].items()}


def _check_grads(P, g, skip=()):
    for k, ref in g["grads"].items():
        if k in skip:
            continue
        assert P[k].grad is not None, f"no grad for {k}"
        assert_close_rel(P[k].grad, ref, 5e-5, f"grad {k}", atol=1e-6)


def test_erp_v4_small():
    g = load_golden("erp_v4_small")
    P = _params(g)
    x = g["inputs"][0].clone().requires_grad_(True)
    conv = om.enhanced_erp_conv_stack(P, "", x)
    assert_close_rel(conv, g["raw"]["conv_stack_out"], TOL, "conv stack")
    y = om.enhanced_erp_encoder(P, "", x, nhead=4)
    assert_close_rel(y, g["outputs"][0], TOL, "encoder output")
    y.backward(torch.from_numpy(g["raw"]["cotangent"]))
    assert_close_rel(x.grad, g["in_grads"][0], 5e-5, "dx")
    _check_grads(P, g)


def test_erp_v4_d128():
    """The same encoder at the BASELINE width (d_model 128, 4 heads of 32, 2 blocks): the fixture the fused
    transformer tail of the CUDA path is held to (tests/test_gpu_modules.py::test_erp_v4_d128_golden_fused_tail)."""
    from multimodal_eeg_fmri_b200.modules import PositionalEncoding
    g = load_golden("erp_v4_d128")
    P = _params(g)
    P["pos_encoder.pe"] = PositionalEncoding(128).pe  # deterministic table, left out of the fixture (2.5 MB)
    x = g["inputs"][0].clone().requires_grad_(True)
    assert_close_rel(om.enhanced_erp_conv_stack(P, "", x), g["raw"]["conv_stack_out"], TOL, "conv stack")
    y = om.enhanced_erp_encoder(P, "", x, nhead=4)
    assert_close_rel(y, g["outputs"][0], TOL, "encoder output")
    y.backward(torch.from_numpy(g["raw"]["cotangent"]))
    assert_close_rel(x.grad, g["in_grads"][0], 5e-5, "dx")
    _check_grads(P, g)


def test_power_v4_small():
    g = load_golden("power_v4_small")
    P = _params(g)
    x = g["inputs"][0].clone().requires_grad_(True)
    y = om.enhanced_power_encoder(P, "", x, nhead=4)
    assert_close_rel(y, g["outputs"][0], TOL, "encoder output")
    y.backward(torch.from_numpy(g["raw"]["cotangent"]))
    assert_close_rel(x.grad, g["in_grads"][0], 5e-5, "dx")
    _check_grads(P, g)


@pytest.mark.parametrize("name", ["lite_erp_small", "lite_pw_small"])
def test_lite_encoders(name):
    g = load_golden(name)
    P = _params(g)
    x = g["inputs"][0].clone().requires_grad_(True)
    y = om.lite_encoder(P, "", x)
    assert_close_rel(y, g["outputs"][0], TOL, "output")
    y.backward(torch.from_numpy(g["raw"]["cotangent"]))
    assert_close_rel(x.grad, g["in_grads"][0], 5e-5, "dx")
    _check_grads(P, g)


def test_trimodal_lite_small():
    g = load_golden("trimodal_lite_small")
    P = _params(g)
    erp, pw, conn = (g["inputs"][i] for i in range(3))
    logits, w, fused = om.trimodal_lite(P, "", erp, pw, conn, conn_boost=1.3)
    assert_close_rel(logits, g["outputs"][0], TOL, "logits")
    assert_close_rel(fused, g["raw"]["fused"], TOL, "fused")
    assert_close_rel(torch.tensor([w["erp_weight"], w["pw_weight"], w["conn_weight"]]), g["raw"]["weights"], TOL, "weights")
    ls = om.label_smoothing_ce(logits, torch.from_numpy(g["raw"]["labels"]))
    assert_close_rel(ls, g["raw"]["ls_ce"], TOL, "label smoothing CE")
    logits.backward(torch.from_numpy(g["raw"]["cotangent"]))
    _check_grads(P, g)


def test_fmri_small():
    g = load_golden("fmri_small")
    P = _params(g)
    out, fused = om.fmri_fusion_net(P, "", g["inputs"][0], g["inputs"][1])
    assert_close_rel(out, g["outputs"][0], TOL, "logits")
    assert_close_rel(fused, g["raw"]["fused"], TOL, "fused")
    out.backward(torch.from_numpy(g["raw"]["cotangent"]))
    _check_grads(P, g)


def test_bridge_small():
    g = load_golden("bridge_small")
    P = _params(g)
    logits, fused, fw, aw = om.bridge_net(P, "", g["inputs"][0], g["inputs"][1], nhead=4)
    assert_close_rel(logits, g["outputs"][0], TOL, "logits")
    assert_close_rel(fused, g["raw"]["fused"], TOL, "fused")
    assert_close_rel(fw, g["raw"]["fusion_weights"], TOL, "fusion weights")
    assert_close_rel(aw, g["raw"]["attn_weights"], TOL, "attention weights")
    e, f = om.bridge_projections(P, "", g["inputs"][0], g["inputs"][1])
    assert_close_rel(e, g["raw"]["eeg_proj"], TOL, "eeg_proj")
    assert_close_rel(f, g["raw"]["fmri_proj"], TOL, "fmri_proj")
    logits.backward(torch.from_numpy(g["raw"]["cotangent"]))
    _check_grads(P, g)


def test_bridge_train_recipe():
    """3 steps of CE -> backward -> clip_grad_norm_(1.0) -> AdamW(1e-4, wd 1e-4) reproduce the reference's
    parameters (_test_bridge.py:775-788,869)."""
    z = load_golden("bridge_train3")["raw"]
    P = {k[4:]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith("sd0/")}
    eeg, fmri, y = (torch.from_numpy(z[k]) for k in ("eeg", "fmri", "labels"))
    state, losses = {}, []
    trainable = [k for k, v in P.items() if v.is_floating_point()]
    for _ in range(3):
        leaves = {k: P[k].clone().requires_grad_(True) for k in trainable}
        logits, *_ = om.bridge_net({**P, **leaves}, "", eeg, fmri, nhead=4)
        loss = torch.nn.functional.cross_entropy(logits, y)
        grads = dict(zip(trainable, torch.autograd.grad(loss, [leaves[k] for k in trainable])))
        om.clip_and_adamw({k: P[k] for k in trainable}, grads, state, lr=1e-4, weight_decay=1e-4, max_norm=1.0)
        losses.append(float(loss))
    assert_close_rel(torch.tensor(losses), z["losses"], 1e-5, "losses")
    for k in trainable:
        a, b = P[k], torch.from_numpy(z["sd3/" + k])
        if k == "cross_attn.in_proj_bias":
            # the key bias has an exactly-zero gradient (softmax is invariant to it); Adam turns the
            # rounding noise of that gradient into +-lr steps of arbitrary sign -> not comparable
            d = a.numel() // 3
            keep = torch.ones_like(a, dtype=torch.bool)
            keep[d:2 * d] = False
            a, b = a[keep], b[keep]
        assert_close_rel(a, b, 1e-5, f"param {k} after 3 steps")


def test_structure_known_answers():
    s = json.loads((GOLDEN / "structure.json").read_text())
    # SURVEY.md section 4 structural KATs
    assert s["bridge_default"]["params"] == 133063
    assert s["erp_v4_64_128"]["params"] == 532800
    assert s["power_v4_64_128"]["params"] == 500032
    assert s["bridge_default"]["keys"]["eeg_proj.0.weight"] == [128, 128]
    assert s["bridge_default"]["keys"]["classifier.4.bias"] == [2]
    w = s["bridge_fusion_weights_init"]
    assert abs(w["eeg_weight"] - 0.5) < 1e-7 and abs(w["temperature"] - 1.0) < 1e-7
