"""Drop-in modules on a B200 against (1) the golden vectors produced by the REAL reference classes
(tests/golden/*.npz, oracle/make_golden.py) and (2) the oracle restatement at larger shapes.
The state_dicts load with strict=True; dropout = 0 / eval where the reference fixture says so.

Tolerance: features, losses and gradients within 1e-3 relative (tf32 tensor-core contractions with
fp32 accumulation and fp32 norms/activations) -- BASELINE.json north_star."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close_rel, assert_zero_grad_noise, bias_before_batchnorm, load_golden

pytestmark = pytest.mark.gpu

# Tolerances: the north-star's 1e-3 relative for features, losses and gradients (tf32 contractions with fp32
# accumulation; the first convs run their forward in the 3-pass fp32-accurate mode, see functional.ConvBnAct).
# Measured on a B200 (XM_PRINT_ERRS=1, profiles/r2_tolerance_audit.txt): every gradient of every fixture <= 8.8e-4.
# Class logits of the 4-6 sample fixtures are small differences of features and keep 3e-3.
TOL = 1e-3
LTOL = 3e-3
GTOL = 1e-3


def _load(module, g):
    module.load_state_dict(g["sd"], strict=True)
    return module.cuda().train()


def _check_grads(module, g, tol=GTOL, skip=()):
    named = dict(module.named_parameters())
    zero = bias_before_batchnorm(module.state_dict().keys()) if module.training else set()
    for k, ref in g["grads"].items():
        if k in skip:
            continue
        assert named[k].grad is not None, f"no grad for {k}"
        if k in zero:  # mathematically zero (train-mode BN follows): only rounding noise on both sides
            assert_zero_grad_noise(named[k].grad, named[k[: -len("bias")] + "weight"].grad, f"grad {k}")
            continue
        assert_close_rel(named[k].grad, ref, tol, f"grad {k}", atol=2e-5)


def _run(module, g, out_index=None):
    ins = [t.cuda().requires_grad_(t.is_floating_point()) for _, t in sorted(g["inputs"].items())]
    out = module(*ins)
    outs = out if isinstance(out, tuple) else (out,)
    outs[0].backward(torch.from_numpy(g["raw"]["cotangent"]).cuda())
    return ins, outs


def test_erp_v4_golden():
    from multimodal_eeg_fmri_b200.enhanced_models_v4 import EnhancedERPEncoder
    g = load_golden("erp_v4_small")
    m = _load(EnhancedERPEncoder(8, 32, 1, 4, 0.0), g)
    conv = m.conv_stack(g["inputs"][0].cuda())
    assert_close_rel(conv.transpose(1, 2), g["raw"]["conv_stack_out"], TOL, "conv stack")
    m.load_state_dict(g["sd"], strict=True)  # reset running stats advanced by the probe above
    ins, outs = _run(m, g)
    assert_close_rel(outs[0], g["outputs"][0], TOL, "encoder output")
    assert_close_rel(ins[0].grad, g["in_grads"][0], GTOL, "dx")
    _check_grads(m, g)
    for k, v in g["sd_after"].items():  # BN running statistics after one train-mode forward
        # running statistics of a tf32 conv output: a few 1e-4 (num_batches_tracked: exact)
        assert_close_rel(m.state_dict()[k], v, 5e-4 if v.is_floating_point() else 0.0, k)


def test_erp_v4_d128_golden_fused_tail():
    """The REAL reference class at the BASELINE width (d_model 128, 4 heads of 32, 2 blocks, L = 48): here the CUDA
    path runs its fused transformer tail (fa::attn_* kernels, ffn::* kernels, resid_ln_*) and the 3-pass first
    convs.  Features and every gradient within 1e-3."""
    from multimodal_eeg_fmri_b200 import functional as XF
    from multimodal_eeg_fmri_b200.enhanced_models_v4 import EnhancedERPEncoder
    g = load_golden("erp_v4_d128")
    m = EnhancedERPEncoder(16, 128, 2, 4, 0.0)
    missing = m.load_state_dict(g["sd"], strict=False)  # the fixture omits only the deterministic PE table
    assert missing.missing_keys == ["pos_encoder.pe"] and not missing.unexpected_keys
    m = m.cuda().train()
    assert XF.transformer_tail_supported(48, 128, 4, "gelu")
    ins, outs = _run(m, g)
    assert_close_rel(outs[0], g["outputs"][0], TOL, "encoder output")
    assert_close_rel(ins[0].grad, g["in_grads"][0], GTOL, "dx")
    named = dict(m.named_parameters())
    tail = {k: v for k, v in g["grads"].items() if not k.startswith("conv_layers.")}
    for k, ref in tail.items():
        assert_close_rel(named[k].grad, ref, TOL, f"grad {k}", atol=2e-5)
    _check_grads(m, g, skip=set(tail))


def test_power_v4_golden():
    from multimodal_eeg_fmri_b200.enhanced_models_v4 import EnhancedPowerEncoder
    g = load_golden("power_v4_small")
    m = _load(EnhancedPowerEncoder(8, 32, 1, 4, 0.0), g)
    ins, outs = _run(m, g)
    assert_close_rel(outs[0], g["outputs"][0], TOL, "encoder output")
    assert_close_rel(ins[0].grad, g["in_grads"][0], GTOL, "dx")
    _check_grads(m, g)


@pytest.mark.parametrize("name,cls", [("lite_erp_small", "LiteERPEncoder"), ("lite_pw_small", "LitePowerEncoder")])
def test_lite_encoders_golden(name, cls):
    from multimodal_eeg_fmri_b200 import crossmodal_v4_enhancements as cm
    g = load_golden(name)
    m = _load(getattr(cm, cls)(8, 24, 0.0), g)
    ins, outs = _run(m, g)
    assert_close_rel(outs[0], g["outputs"][0], TOL, "output")
    assert_close_rel(ins[0].grad, g["in_grads"][0], GTOL, "dx")
    _check_grads(m, g)


def test_trimodal_lite_golden():
    from multimodal_eeg_fmri_b200.crossmodal_v4_enhancements import EnhancedTriModalFusionNetV4Lite, LabelSmoothingCrossEntropy
    g = load_golden("trimodal_lite_small")
    m = _load(EnhancedTriModalFusionNetV4Lite(8, 8, 30, hidden_dim=24, num_classes=2, dropout=0.0, conn_boost=1.3), g)
    erp, pw, conn = (g["inputs"][i].cuda() for i in range(3))
    logits, weights, fused = m(erp, pw, conn, return_fusion_weights=True, return_fused_feats=True)
    assert_close_rel(logits, g["outputs"][0], LTOL, "logits")
    assert_close_rel(fused, g["raw"]["fused"], TOL, "fused")
    w = torch.tensor([weights["erp_weight"], weights["pw_weight"], weights["conn_weight"]])
    assert_close_rel(w, g["raw"]["weights"], TOL, "fusion weights")
    ls = LabelSmoothingCrossEntropy(0.1)(logits, torch.from_numpy(g["raw"]["labels"]).cuda())
    assert_close_rel(ls, g["raw"]["ls_ce"], TOL, "label-smoothing CE (loss)")
    m.load_state_dict(g["sd"], strict=True)
    m.zero_grad()
    _run(m, g)
    _check_grads(m, g)


def test_fmri_golden():
    from multimodal_eeg_fmri_b200.fmri_utils import fMRIFusionNet
    g = load_golden("fmri_small")
    m = _load(fMRIFusionNet(20, 50, hidden_dim=16, num_classes=2, dropout=0.0), g)
    out, fused = m(g["inputs"][0].cuda(), g["inputs"][1].cuda(), return_features=True)
    assert_close_rel(out, g["outputs"][0], LTOL, "logits")
    assert_close_rel(fused, g["raw"]["fused"], TOL, "fused")
    m.load_state_dict(g["sd"], strict=True)
    m.zero_grad()
    _run(m, g)
    _check_grads(m, g)


def test_bridge_golden():
    from multimodal_eeg_fmri_b200.bridge_utils import EEGfMRIBridgeFusionNet
    g = load_golden("bridge_small")
    m = EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().eval()  # fixture taken in eval(): gate_net's hard-coded Dropout(0.2)
    eeg, fmri = g["inputs"][0].cuda(), g["inputs"][1].cuda()
    logits, fused, fw, aw = m(eeg, fmri, return_features=True, return_weights=True)
    assert logits.shape == (5, 2) and fused.shape == (5, 32) and fw.shape == (5, 2) and aw.shape == (5, 1, 2)
    assert_close_rel(logits, g["outputs"][0], TOL, "logits")
    assert_close_rel(fused, g["raw"]["fused"], TOL, "fused")
    assert_close_rel(fw, g["raw"]["fusion_weights"], TOL, "fusion weights")
    assert_close_rel(aw, g["raw"]["attn_weights"], TOL, "attention weights")
    e, f = m.project(eeg, fmri)
    assert_close_rel(e, g["raw"]["eeg_proj"], TOL, "eeg_proj")
    assert_close_rel(f, g["raw"]["fmri_proj"], TOL, "fmri_proj")
    m.zero_grad()
    ins = [eeg.clone().requires_grad_(True), fmri.clone().requires_grad_(True)]
    m(*ins).backward(torch.from_numpy(g["raw"]["cotangent"]).cuda())
    assert_close_rel(ins[0].grad, g["in_grads"][0], GTOL, "d eeg")
    assert_close_rel(ins[1].grad, g["in_grads"][1], GTOL, "d fmri")
    # the key bias of the 1x2 attention has an exactly-zero gradient (softmax shift invariance)
    _check_grads(m, g, skip=("cross_attn.in_proj_bias",))
    d = 32
    gb = dict(m.named_parameters())["cross_attn.in_proj_bias"].grad.cpu()
    ref = g["grads"]["cross_attn.in_proj_bias"]
    keep = torch.ones(3 * d, dtype=torch.bool)
    keep[d:2 * d] = False
    assert_close_rel(gb[keep], ref[keep], GTOL, "in_proj_bias (q, v parts)")
    assert float(gb[~keep].abs().max()) < 1e-5


def test_bridge_attribution_helpers_golden():
    """BridgeGradientSaliency / BridgeIntegratedGradients (all interpolation points in one batch) /
    extract_attention_and_fusion_weights against the reference classes' outputs (bridge_utils.py:158-270)."""
    from test_host_pipeline_cpu import check_bridge_xai
    check_bridge_xai("cuda", GTOL)


def test_bridge_train_recipe_golden():
    """3 steps of CE -> backward -> clip_grad_norm_(1.0) -> AdamW(1e-4, wd 1e-4) through
    train_bridge_epoch reproduce the reference's losses and parameters (_test_bridge.py:775-788,869)."""
    from multimodal_eeg_fmri_b200.bridge_utils import EEGfMRIBridgeFusionNet
    from multimodal_eeg_fmri_b200.training import train_bridge_epoch
    z = load_golden("bridge_train3")["raw"]
    m = EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    m.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}, strict=True)
    m = m.cuda()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
    loader = [(torch.from_numpy(z["eeg"]), torch.from_numpy(z["fmri"]), torch.from_numpy(z["labels"]), list(range(8)))]

    class _Eval(torch.nn.Module):  # the fixture ran the model in eval() (see oracle/make_golden.py)
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def train(self, mode=True):
            return super().train(False)

        def forward(self, a, b):
            return self.inner(a, b)

    wrapped = _Eval(m).eval()
    losses = [train_bridge_epoch(wrapped, loader, opt, torch.nn.CrossEntropyLoss(), "cuda", 1.0) for _ in range(3)]
    assert_close_rel(torch.tensor(losses), z["losses"], TOL, "losses")
    sd = m.state_dict()
    for k in z.files:
        if not k.startswith("sd3/") or k[4:] == "cross_attn.in_proj_bias":
            continue
        # three AdamW steps move each weight by <= 3e-4; compare the UPDATE, not just the value
        ref0, ref3 = torch.from_numpy(z["sd0/" + k[4:]]), torch.from_numpy(z[k])
        assert_close_rel(sd[k[4:]], ref3, 1e-4, f"param {k[4:]} after 3 steps", atol=3e-5)
        if ref0.numel() > 64:
            assert_close_rel(sd[k[4:]].cpu() - ref0, ref3 - ref0, 0.12, f"update of {k[4:]}", atol=2e-5)


# ------------------------------------------------------------------ larger shapes vs the oracle restatement
def _sd_cpu(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def test_erp_v4_config1_shape_vs_oracle():
    """EnhancedERPEncoder(64, 128, 2, 4) on (8, 64, 500): the BASELINE config-1 sample shape."""
    from multimodal_eeg_fmri_b200.enhanced_models_v4 import EnhancedERPEncoder
    from oracle import models as om
    torch.manual_seed(42)
    m = EnhancedERPEncoder(64, 128, 2, 4, 0.0)
    P = _sd_cpu(m)
    x = torch.randn(8, 64, 500)
    cot = torch.randn(8, 128)
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    y.backward(cot.cuda())
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")}
    xo = x.clone().requires_grad_(True)
    yo = om.enhanced_erp_encoder({**P, **leaves}, "", xo, nhead=4)
    yo.backward(cot)
    assert_close_rel(y, yo, TOL, "encoder output")
    # input gradient: the full depth of the encoder (3 conv + 2 transformer blocks, ~25 chained contractions)
    assert_close_rel(xg.grad, xo.grad, GTOL, "dx")
    zero = bias_before_batchnorm(m.state_dict().keys())
    for k, p in m.named_parameters():
        if k in zero:
            assert_zero_grad_noise(p.grad, dict(m.named_parameters())[k[: -len("bias")] + "weight"].grad, f"grad {k}")
        else:
            assert_close_rel(p.grad, leaves[k].grad, GTOL, f"grad {k}", atol=2e-5)


def test_power_v4_full_length_vs_oracle():
    """EnhancedPowerEncoder(64, 128, 2, 4) on (4, 64, 500): no pooling in front of the transformer, so the attention
    runs over L = 500 tokens -- the fused core's four-block path (enhanced_models_v4.py:196-285)."""
    from multimodal_eeg_fmri_b200.enhanced_models_v4 import EnhancedPowerEncoder
    from multimodal_eeg_fmri_b200 import functional as XF
    from oracle import models as om
    assert XF.transformer_tail_supported(500, 128, 4, "gelu")
    torch.manual_seed(43)
    m = EnhancedPowerEncoder(64, 128, 2, 4, 0.0)
    P = _sd_cpu(m)
    x = torch.randn(4, 64, 500)
    cot = torch.randn(4, 128)
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    y.backward(cot.cuda())
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")}
    xo = x.clone().requires_grad_(True)
    yo = om.enhanced_power_encoder({**P, **leaves}, "", xo, nhead=4)
    yo.backward(cot)
    assert_close_rel(y, yo, TOL, "encoder output")
    assert_close_rel(xg.grad, xo.grad, GTOL, "dx")
    zero = bias_before_batchnorm(m.state_dict().keys())
    for k, p in m.named_parameters():
        if k in zero:
            assert_zero_grad_noise(p.grad, dict(m.named_parameters())[k[: -len("bias")] + "weight"].grad, f"grad {k}")
        else:
            assert_close_rel(p.grad, leaves[k].grad, GTOL, f"grad {k}", atol=2e-5)


def test_fmri_config2_shape_vs_oracle():
    """fMRIFusionNet(400, 40000) on batch 64 with the ROI aggregation on the device (config 2)."""
    from multimodal_eeg_fmri_b200 import fmri_utils
    from oracle import models as om
    torch.manual_seed(42)
    m = fmri_utils.fMRIFusionNet(400, 40000, 64, 2, 0.0)
    P = _sd_cpu(m)
    roi = torch.randn(64, 100, 200)
    conn = torch.randn(64, 40000)
    m = m.cuda().train()
    act = fmri_utils.aggregate_roi_timeseries(roi.cuda(), "both")
    assert_close_rel(act, om.roi_meanstd(roi), 1e-5, "ROI mean/std")
    out, fused = m(act, conn.cuda(), return_features=True)
    oo, of = om.fmri_fusion_net(P, "", om.roi_meanstd(roi), conn)
    assert_close_rel(out, oo, LTOL, "logits")
    assert_close_rel(fused, of, TOL, "fused")
    assert m.get_fusion_weights() == pytest.approx({"activation": 0.5, "connectivity": 0.5})


def test_entry_points_run():
    from multimodal_eeg_fmri_b200 import run_fmri_v11, run_training_lite
    l1 = run_training_lite.main(steps=3, batch=8, channels=16, samples=64, conn_dim=40)
    l2 = run_fmri_v11.main(steps=3, batch=16, n_roi=12, n_tr=20)
    assert all(np.isfinite(l1)) and all(np.isfinite(l2))


def test_dropout_train_mode_is_statistically_sane():
    """Train-mode dropout cannot be bit-matched to torch's RNG (SURVEY.md section 7 hard part 3): check that
    the expected output equals the dropout-free output (inverted-dropout scaling) on a linear probe."""
    from multimodal_eeg_fmri_b200 import functional as XF
    x = torch.randn(4096, 256, device="cuda")
    XF.manual_seed(123)
    outs = torch.stack([XF.act_dropout(x, "none", 0.3, True) for _ in range(64)]).mean(0)
    assert float((outs - x).abs().mean() / x.abs().mean()) < 0.12
    XF.manual_seed(123)
    a = XF.act_dropout(x, "relu", 0.3, True)
    XF.manual_seed(123)
    b = XF.act_dropout(x, "relu", 0.3, True)
    assert torch.equal(a, b), "same seed -> same mask"
    assert torch.equal(XF.act_dropout(x, "relu", 0.3, False), F.relu(x))
