"""CPU oracle (test infrastructure): functional restatement of the reference modules on the
paired-step hot path.  Each function takes a reference `state_dict` (same key names as the
reference classes) plus a key prefix and evaluates the module with plain torch.nn.functional
calls -- no nn.Module, no parameters of its own -- so the same tensors can be fed to the CUDA
modules and to this oracle.  Dropout is the identity here (parity runs use dropout=0 / eval, SURVEY.md
section 7 hard part 3); BatchNorm uses batch statistics when `train=True`.

Pinned against the real reference classes by oracle/make_golden.py -> tests/golden/*.npz and
tests/test_oracle_golden.py.  Citations are reference file:line.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# Optional operand-rounding model, `fn(key_prefix, tensor) -> tensor`, applied to the operands of every
# matmul/conv below.  None (the default, and what every parity test uses) is the reference's plain
# arithmetic.  tools/full_scale_parity.py sets it to round-to-nearest tf32 on the EEG encoder to measure the
# error floor ANY single-pass tf32 implementation of that encoder has (it is not part of the oracle proper).
OPERAND_ROUNDING = None


def _r(pre: str, t):
    return t if OPERAND_ROUNDING is None else OPERAND_ROUNDING(pre, t)


def _bn(P: SD, pre: str, x, train: bool, eps: float = 1e-5):
    # nn.BatchNorm1d: batch statistics in train mode, running statistics in eval mode
    if train:
        return F.batch_norm(x, None, None, P[pre + "weight"], P[pre + "bias"], True, 0.1, eps)
    return F.batch_norm(x, P[pre + "running_mean"], P[pre + "running_var"], P[pre + "weight"], P[pre + "bias"], False, 0.1, eps)


def _lin(P: SD, pre: str, x):
    return F.linear(_r(pre, x), _r(pre, P[pre + "weight"]), P.get(pre + "bias"))


def _conv(P: SD, pre: str, x):
    w = P[pre + "weight"]
    return F.conv1d(_r(pre, x), _r(pre, w), P.get(pre + "bias"), padding=w.shape[-1] // 2)


def _ln(P: SD, pre: str, x, eps: float = 1e-5):
    w = P[pre + "weight"]
    return F.layer_norm(x, (w.shape[0],), w, P[pre + "bias"], eps)


# --------------------------------------------------------------------------- transformer tail
def positional_encoding(P: SD, pre: str, x):
    """EEG_CODE/enhanced_models_v4.py:30-55 (crossmodal_v4_enhancements.py:29-51), batch-first branch:
    x (B, L, d) + pe[:L] broadcast over the batch."""
    pe = P[pre + "pe"]  # (max_len, 1, d)
    if x.dim() == 3 and x.size(1) != 1:
        return x + pe[: x.size(1), 0, :].unsqueeze(0)
    return x + pe[: x.size(0)]


def multihead_attention(P: SD, pre: str, q, k, v, nhead: int, need_weights: bool = False):
    """nn.MultiheadAttention(batch_first=True) as called at enhanced_models_v4.py:71-73,98 and
    bridge_utils.py:48,80-82: packed in_proj, scaled dot-product per head, out_proj; the returned
    weights are averaged over heads (need_weights default)."""
    d = q.shape[-1]
    W, b = _r(pre, P[pre + "in_proj_weight"]), P[pre + "in_proj_bias"]
    q, k, v = _r(pre, q), _r(pre, k), _r(pre, v)
    qp = F.linear(q, W[:d], b[:d])
    kp = F.linear(k, W[d:2 * d], b[d:2 * d])
    vp = F.linear(v, W[2 * d:], b[2 * d:])
    B, Lq, _ = qp.shape
    Lk = kp.shape[1]
    dh = d // nhead
    qh = qp.view(B, Lq, nhead, dh).transpose(1, 2)
    kh = kp.view(B, Lk, nhead, dh).transpose(1, 2)
    vh = vp.view(B, Lk, nhead, dh).transpose(1, 2)
    att = torch.softmax(_r(pre, qh) @ _r(pre, kh).transpose(-1, -2) / math.sqrt(dh), dim=-1)
    out = (_r(pre, att) @ _r(pre, vh)).transpose(1, 2).reshape(B, Lq, d)
    out = F.linear(_r(pre, out), _r(pre, P[pre + "out_proj.weight"]), P[pre + "out_proj.bias"])
    return (out, att.mean(1)) if need_weights else (out, None)


def transformer_block(P: SD, pre: str, x, nhead: int):
    """TemporalTransformerBlock.forward, enhanced_models_v4.py:89-107: pre-norm attention and FFN
    residual branches (GELU)."""
    h = _ln(P, pre + "norm1.", x)
    h, _ = multihead_attention(P, pre + "self_attn.", h, h, h, nhead)
    x = x + h
    h = _ln(P, pre + "norm2.", x)
    h = _lin(P, pre + "linear2.", F.gelu(_lin(P, pre + "linear1.", h)))
    return x + h


def _transformer_tail(P: SD, pre: str, x, nhead: int):
    x = positional_encoding(P, pre + "pos_encoder.", x.transpose(1, 2))
    i = 0
    while f"{pre}transformer_layers.{i}.norm1.weight" in P:
        x = transformer_block(P, f"{pre}transformer_layers.{i}.", x, nhead)
        i += 1
    x = x.transpose(1, 2).mean(-1)  # AdaptiveAvgPool1d(1) + Flatten
    return F.gelu(_lin(P, pre + "output_proj.2.", x))


# --------------------------------------------------------------------------- EEG encoders
def enhanced_erp_encoder(P: SD, pre: str, x, nhead: int = 4, train: bool = True):
    """EnhancedERPEncoder, enhanced_models_v4.py:114-193 (== crossmodal_v4_enhancements.py:93-143):
    Conv(k7)-BN-GELU, Conv(k5)-BN-GELU-MaxPool2, Conv(k3)-BN-GELU, PE + transformer blocks, mean-pool,
    Linear-GELU."""
    c = pre + "conv_layers."
    x = F.gelu(_bn(P, c + "1.", _conv(P, c + "0.", x), train))
    x = F.max_pool1d(F.gelu(_bn(P, c + "5.", _conv(P, c + "4.", x), train)), 2)
    x = F.gelu(_bn(P, c + "10.", _conv(P, c + "9.", x), train))
    return _transformer_tail(P, pre, x, nhead)


def enhanced_erp_conv_stack(P: SD, pre: str, x, train: bool = True):
    """Only the conv_layers Sequential of EnhancedERPEncoder (enhanced_models_v4.py:128-144)."""
    c = pre + "conv_layers."
    x = F.gelu(_bn(P, c + "1.", _conv(P, c + "0.", x), train))
    x = F.max_pool1d(F.gelu(_bn(P, c + "5.", _conv(P, c + "4.", x), train)), 2)
    return F.gelu(_bn(P, c + "10.", _conv(P, c + "9.", x), train))


def enhanced_power_encoder(P: SD, pre: str, x, nhead: int = 4, train: bool = True):
    """EnhancedPowerEncoder, enhanced_models_v4.py:196-285: three parallel Conv(k3/5/7)-BN-GELU
    scales, concat, 1x1 Conv-BN-GELU, transformer tail without pooling."""
    s = [F.gelu(_bn(P, f"{pre}conv_scale{i}.1.", _conv(P, f"{pre}conv_scale{i}.0.", x), train)) for i in (1, 2, 3)]
    x = torch.cat(s, dim=1)
    x = F.gelu(_bn(P, pre + "fusion.1.", _conv(P, pre + "fusion.0.", x), train))
    return _transformer_tail(P, pre, x, nhead)


def lite_encoder(P: SD, pre: str, x, train: bool = True):
    """LiteERPEncoder / LitePowerEncoder, crossmodal_v4_enhancements.py:817-877:
    Conv-BN-GELU-(Drop)-MaxPool2, Conv-BN-GELU-(Drop)-AvgPool(1), Flatten-Linear-GELU."""
    c = pre + "conv_layers."
    x = F.max_pool1d(F.gelu(_bn(P, c + "1.", _conv(P, c + "0.", x), train)), 2)
    x = F.gelu(_bn(P, c + "6.", _conv(P, c + "5.", x), train)).mean(-1)
    return F.gelu(_lin(P, pre + "output.1.", x))


def enhanced_conn_encoder(P: SD, pre: str, x, train: bool = True):
    """EnhancedConnEncoder, crossmodal_v4_enhancements.py:684-739."""
    if x.dim() > 2:
        x = x.reshape(x.size(0), -1)
    x = F.gelu(_bn(P, pre + "proj1.1.", _lin(P, pre + "proj1.0.", x), train))
    x = F.gelu(_bn(P, pre + "proj2.1.", _lin(P, pre + "proj2.0.", x), train))
    a = torch.sigmoid(_lin(P, pre + "attention.2.", torch.tanh(_lin(P, pre + "attention.0.", x))))
    x = x * a
    return F.gelu(_bn(P, pre + "output.1.", _lin(P, pre + "output.0.", x), train))


def hybrid_fusion(P: SD, pre: str, erp, pw, conn, conn_boost: float, train: bool = True):
    """HybridFusionModule.forward, crossmodal_v4_enhancements.py:778-810.  Returns (fused, weights)
    with the same three weight entries the reference reports."""
    g = _lin(P, pre + "erp_pw_gate.3.", F.gelu(_lin(P, pre + "erp_pw_gate.0.", torch.cat([erp, pw], 1))))
    g = torch.softmax(g, dim=-1)
    early = g[:, 0:1] * erp + g[:, 1:2] * pw
    fw = torch.softmax(P[pre + "final_gate"], dim=0)
    fused = F.gelu(_bn(P, pre + "late_fusion.1.", _lin(P, pre + "late_fusion.0.", torch.cat([early, conn * conn_boost], 1)), train))
    weights = {
        "erp_weight": float(g[:, 0].mean()) * float(fw[0]),
        "pw_weight": float(g[:, 1].mean()) * float(fw[0]),
        "conn_weight": float(fw[1]) * conn_boost,
    }
    return fused, weights


def trimodal_lite(P: SD, pre: str, erp, pw, conn, conn_boost: float = 1.3, train: bool = True):
    """EnhancedTriModalFusionNetV4Lite.forward, crossmodal_v4_enhancements.py:920-944.
    Returns (logits, weights, fused)."""
    e = lite_encoder(P, pre + "erp_encoder.", erp, train)
    p = lite_encoder(P, pre + "pw_encoder.", pw, train)
    c = enhanced_conn_encoder(P, pre + "conn_encoder.", conn, train)
    fused, w = hybrid_fusion(P, pre + "fusion.", e, p, c, conn_boost, train)
    h = F.gelu(_bn(P, pre + "classifier.1.", _lin(P, pre + "classifier.0.", fused), train))
    return _lin(P, pre + "classifier.4.", h), w, fused


def label_smoothing_ce(pred, target, smoothing: float = 0.1):
    """LabelSmoothingCrossEntropy.forward, crossmodal_v4_enhancements.py:672-677."""
    logp = F.log_softmax(pred, dim=-1)
    nll = -logp.gather(-1, target.unsqueeze(1)).squeeze(1)
    return ((1.0 - smoothing) * nll + smoothing * (-logp.mean(-1))).mean()


# --------------------------------------------------------------------------- fMRI
def roi_meanstd(x: torch.Tensor) -> torch.Tensor:
    """fMRI_CODE/fmri_utils.py:140-147 (agg_method='both') on a batch: x (B, TR, ROI) ->
    (B, 2*ROI) = concat(mean over TR, population std over TR) after nan_to_num."""
    x = torch.nan_to_num(x, nan=0.0)
    return torch.cat([x.mean(1), x.std(1, unbiased=False)], dim=1)


def roi_connectivity(x: torch.Tensor) -> torch.Tensor:
    """(B, TR, ROI) -> (B, ROI*ROI): flattened per-sample numpy.corrcoef of the ROI columns (NaN -> 0 first).
    AUTHORED definition, PARITY UNPINNED: the reference only reads precomputed connectivity matrices from CSV
    (fmri_utils.py:161-198); SURVEY.md section 8d defines the synthetic connectivity input as this corrcoef.
    Pinned by a numpy.corrcoef known-answer test (tests/test_oracle_paired_step.py)."""
    xc = torch.nan_to_num(x)
    xc = xc - xc.mean(1, keepdim=True)
    c = xc.transpose(1, 2) @ xc
    d = torch.sqrt(torch.diagonal(c, dim1=1, dim2=2))
    return (c / d[:, :, None] / d[:, None, :]).clamp(-1, 1).reshape(x.shape[0], -1)


def fmri_mlp_encoder(P: SD, pre: str, x, train: bool = True):
    """ActivationEncoder / ConnectivityEncoder, fMRI_CODE/fmri_utils.py:23-56."""
    e = pre + "encoder."
    x = F.relu(_bn(P, e + "1.", _lin(P, e + "0.", x), train))
    return F.relu(_bn(P, e + "5.", _lin(P, e + "4.", x), train))


def fmri_fusion_net(P: SD, pre: str, activation, connectivity, train: bool = True, task: str = "classification"):
    """fMRIFusionNet.forward, fMRI_CODE/fmri_utils.py:90-103.  Returns (output, fused)."""
    a = fmri_mlp_encoder(P, pre + "activation_encoder.", activation, train)
    c = fmri_mlp_encoder(P, pre + "connectivity_encoder.", connectivity, train)
    w = torch.softmax(torch.stack([P[pre + "activation_weight"], P[pre + "connectivity_weight"]]), dim=0)
    comb = torch.cat([a * w[0], c * w[1]], dim=1)
    fused = F.relu(_bn(P, pre + "fusion.1.", _lin(P, pre + "fusion.0.", comb), train))
    out = _lin(P, pre + "head.3.", F.relu(_lin(P, pre + "head.0.", fused)))
    if task == "regression":
        out = out.squeeze(-1)
    return out, fused


# --------------------------------------------------------------------------- bridge
def learned_fusion(P: SD, pre: str, feats):
    """LearnedFusionModule.forward, crossmodal_v4_enhancements.py:241-271. Returns (fused, weights)."""
    temp = P[pre + "temperature"]
    static = torch.softmax(P[pre + "fusion_logits"] / temp, dim=0)
    dyn = _lin(P, pre + "gate_net.3.", F.gelu(_lin(P, pre + "gate_net.0.", torch.cat(feats, 1))))
    dyn = torch.softmax(dyn / temp, dim=1)
    w = 0.5 * static.unsqueeze(0) + 0.5 * dyn
    fused = (torch.stack(feats, 1) * w.unsqueeze(2)).sum(1)
    return fused, w


def bridge_projections(P: SD, pre: str, eeg, fmri):
    """The two shared-space projections of EEGfMRIBridgeFusionNet (bridge_utils.py:34-45,71-72):
    Linear-LayerNorm-GELU each.  These are the InfoNCE embeddings."""
    e = F.gelu(_ln(P, pre + "eeg_proj.1.", _lin(P, pre + "eeg_proj.0.", eeg)))
    f = F.gelu(_ln(P, pre + "fmri_proj.1.", _lin(P, pre + "fmri_proj.0.", fmri)))
    return e, f


def bridge_net(P: SD, pre: str, eeg, fmri, nhead: int = 4):
    """EEGfMRIBridgeFusionNet.forward, bridge_utils.py:68-101.
    Returns (logits, fused, fusion_weights (B,2), attn_weights (B,1,2))."""
    e, f = bridge_projections(P, pre, eeg, fmri)
    seq = torch.stack([e, f], dim=1)
    att, aw = multihead_attention(P, pre + "cross_attn.", e.unsqueeze(1), seq, seq, nhead, need_weights=True)
    fused, fw = learned_fusion(P, pre + "fusion.", [att.squeeze(1), f])
    h = F.relu(_ln(P, pre + "classifier.1.", _lin(P, pre + "classifier.0.", fused)))
    return _lin(P, pre + "classifier.4.", h), fused, fw, aw


# --------------------------------------------------------------------------- train-step recipe
def clip_and_adamw(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Dict[str, dict], lr: float,
                   weight_decay: float, max_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8):
    """`clip_grad_norm_(max_norm)` then `AdamW.step()` as in _test_bridge.py:784-786,
    fMRI_CODE/run_fmri_v11.py:446-448, EEG_CODE/run_training_lite.py:487-488.
    Mutates `params` / `state` in place; returns the pre-clip total gradient norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for k, p in params.items():
        if k not in grads:
            continue
        g = grads[k] * coef
        st = state.setdefault(k, {"step": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["step"] += 1
        t = st["step"]
        p.mul_(1 - lr * weight_decay)
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st["v"].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        p.addcdiv_(st["m"], denom, value=-lr / (1 - b1 ** t))
    return total
