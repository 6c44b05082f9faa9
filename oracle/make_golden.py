"""Generate tests/golden/*.npz from the REAL reference classes (run in the dev container only).

    python oracle/make_golden.py            # needs /root/reference (read-only mount)

The reference is Python, so it cannot travel to the GPU box: this script imports its classes,
runs them on seeded inputs (dropout = 0, train-mode BatchNorm, torch.manual_seed(42)) and commits
small input / state_dict / output / gradient fixtures.  The oracle restatement (oracle/models.py)
and the CUDA modules are both tested against these files.  Test infrastructure only.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path(os.environ.get("XM_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    """Import the reference modules (torch_geometric / optuna stubbed: only touched at import,
    SURVEY.md section 8c).  Banners printed at import are swallowed."""
    if not REF.exists():
        raise SystemExit(f"{REF} not found: golden vectors can only be regenerated in the dev container")
    sys.path.insert(0, str(REF))
    for name in ("torch_geometric", "torch_geometric.nn", "torch_geometric.data", "optuna"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["torch_geometric.nn"].GATv2Conv = object
    sys.modules["torch_geometric.nn"].global_mean_pool = object
    sys.modules["torch_geometric.data"].Data = object
    sys.modules["torch_geometric.data"].Batch = object
    sys.modules["optuna"].Trial = object
    sys.modules["torch_geometric"].nn = sys.modules["torch_geometric.nn"]
    sys.modules["torch_geometric"].data = sys.modules["torch_geometric.data"]
    with contextlib.redirect_stdout(io.StringIO()):
        import bridge_utils  # noqa: F401
        from EEG_CODE import crossmodal_v4_enhancements as cm
        from EEG_CODE import enhanced_models_v4 as em
        from fMRI_CODE import fmri_utils as fu
    return cm, em, fu, sys.modules["bridge_utils"]


def _np(t):
    return t.detach().cpu().numpy().copy()  # copy: .numpy() aliases the (later updated) parameter


def _pack(store: dict, prefix: str, sd):
    for k, v in sd.items():
        store[f"{prefix}/{k}"] = _np(v)


def _run(module, inputs, out_index=None, seed=0, train=True):
    """forward + backward with a seeded cotangent; returns outputs (tuple) and {param/input: grad}."""
    module.train(train)
    ins = [i.clone().requires_grad_(i.is_floating_point()) for i in inputs]
    out = module(*ins)
    outs = out if isinstance(out, tuple) else (out,)
    main = outs[0] if out_index is None else outs[out_index]
    g = torch.Generator().manual_seed(1000 + seed)
    cot = torch.randn(main.shape, generator=g)
    main.backward(cot)
    grads = {n: p.grad for n, p in module.named_parameters() if p.grad is not None}
    in_grads = [i.grad for i in ins]
    return outs, cot, grads, in_grads


def case_module(name, module, inputs, extra=None, seed=0, train=True):
    store = {}
    sd0 = {k: v.clone() for k, v in module.state_dict().items()}
    outs, cot, grads, in_grads = _run(module, inputs, seed=seed, train=train)
    _pack(store, "sd", sd0)
    # BN running stats after one train-mode forward
    _pack(store, "sd_after", {k: v for k, v in module.state_dict().items() if "running_" in k or "num_batches" in k})
    for i, x in enumerate(inputs):
        store[f"in/{i}"] = _np(x)
        if in_grads[i] is not None:
            store[f"in_grad/{i}"] = _np(in_grads[i])
    for i, o in enumerate(outs):
        if torch.is_tensor(o):
            store[f"out/{i}"] = _np(o)
    store["cotangent"] = _np(cot)
    _pack(store, "grad", grads)
    if extra:
        store.update(extra)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **store)
    print(f"wrote {name}.npz ({len(store)} arrays)")


def main():
    cm, em, fu, bu = import_reference()
    torch.manual_seed(42)
    g = torch.Generator().manual_seed(42)
    rn = lambda *s: torch.randn(*s, generator=g)

    # --- v4 encoders (crossmodal copy == enhanced_models_v4 copy, verified below)
    erp = cm.EnhancedERPEncoder(8, 32, 1, 4, 0.0)
    x = rn(4, 8, 64)
    erp.train()
    conv_out = erp.conv_layers(x)
    case_module("erp_v4_small", erp, [x], extra={"conv_stack_out": _np(conv_out)}, seed=1)

    torch.manual_seed(42)
    erp_em = em.EnhancedERPEncoder(8, 32, 1, 4, 0.0)
    torch.manual_seed(42)
    erp_cm = cm.EnhancedERPEncoder(8, 32, 1, 4, 0.0)
    erp_em.train(); erp_cm.train()
    assert float((erp_em(x) - erp_cm(x)).abs().max()) == 0.0, "enhanced_models_v4 and crossmodal copies diverge"

    torch.manual_seed(43)
    pw = cm.EnhancedPowerEncoder(8, 32, 1, 4, 0.0)
    case_module("power_v4_small", pw, [rn(4, 8, 48)], seed=2)

    torch.manual_seed(44)
    case_module("lite_erp_small", cm.LiteERPEncoder(8, 24, 0.0), [rn(4, 8, 64)], seed=3)
    torch.manual_seed(45)
    case_module("lite_pw_small", cm.LitePowerEncoder(8, 24, 0.0), [rn(4, 8, 64)], seed=4)

    # --- tri-modal lite (config 1 structure, small)
    torch.manual_seed(46)
    tri = cm.EnhancedTriModalFusionNetV4Lite(8, 8, 30, hidden_dim=24, num_classes=2, dropout=0.0, conn_boost=1.3)
    erp_in, pw_in, conn_in = rn(6, 8, 64), rn(6, 8, 64), rn(6, 30)
    tri.train()
    logits, weights, fused = tri(erp_in, pw_in, conn_in, return_fusion_weights=True, return_fused_feats=True)
    y = torch.tensor([0, 1, 1, 0, 1, 0])
    ls = cm.LabelSmoothingCrossEntropy(0.1)(logits, y)
    tri.zero_grad()
    case_module("trimodal_lite_small", tri, [erp_in, pw_in, conn_in],
                extra={"fused": _np(fused), "labels": y.numpy(), "ls_ce": _np(ls),
                       "weights": np.array([weights["erp_weight"], weights["pw_weight"], weights["conn_weight"]])},
                seed=5)

    # --- fMRI
    torch.manual_seed(47)
    fm = fu.fMRIFusionNet(20, 50, hidden_dim=16, num_classes=2, dropout=0.0)
    act, conn = rn(6, 20), rn(6, 50)
    fm.train()
    _, fused = fm(act, conn, return_features=True)
    case_module("fmri_small", fm, [act, conn], extra={"fused": _np(fused)}, seed=6)

    # --- bridge.  LearnedFusionModule.gate_net carries a hard-coded nn.Dropout(0.2)
    # (crossmodal_v4_enhancements.py:236) that the constructor's dropout=0 does not reach, so the
    # bridge fixtures are taken in eval() mode (the model has no BatchNorm: eval == train minus dropout).
    torch.manual_seed(48)
    br = bu.EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    eeg, fmri = rn(5, 32), rn(5, 16)
    br.eval()
    lg, fz, fw, aw = br(eeg, fmri, return_features=True, return_weights=True)
    e_proj, f_proj = br.eeg_proj(eeg), br.fmri_proj(fmri)
    case_module("bridge_small", br, [eeg, fmri],
                extra={"fused": _np(fz), "fusion_weights": _np(fw), "attn_weights": _np(aw),
                       "eeg_proj": _np(e_proj), "fmri_proj": _np(f_proj)}, seed=7, train=False)

    # --- train-step recipe: 3 steps of CE + clip_grad_norm_(1.0) + AdamW(1e-4, wd 1e-4)  (_test_bridge.py:775-788,869)
    torch.manual_seed(49)
    br = bu.EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    store = {}
    _pack(store, "sd0", br.state_dict())
    opt = torch.optim.AdamW(br.parameters(), lr=1e-4, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss()
    eeg, fmri, y = rn(8, 32), rn(8, 16), torch.tensor([0, 1, 0, 1, 1, 0, 0, 1])
    losses = []
    br.eval()  # see the gate_net dropout note above
    for _ in range(3):
        opt.zero_grad()
        loss = crit(br(eeg, fmri), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(br.parameters(), 1.0)
        opt.step()
        losses.append(float(loss))
    _pack(store, "sd3", br.state_dict())
    store.update({"eeg": _np(eeg), "fmri": _np(fmri), "labels": y.numpy(), "losses": np.array(losses)})
    np.savez_compressed(OUT / "bridge_train3.npz", **store)
    print("wrote bridge_train3.npz")

    # --- structural known answers at the BASELINE shapes
    def nparams(m):
        return sum(p.numel() for p in m.parameters())

    def keyshapes(m):
        return {k: list(v.shape) for k, v in m.state_dict().items()}

    with contextlib.redirect_stdout(io.StringIO()):
        struct = {
            "bridge_default": {"params": nparams(bu.EEGfMRIBridgeFusionNet()), "keys": keyshapes(bu.EEGfMRIBridgeFusionNet())},
            "erp_v4_64_128": {"params": nparams(cm.EnhancedERPEncoder(64, 128, 2, 4)), "keys": keyshapes(cm.EnhancedERPEncoder(64, 128, 2, 4))},
            "power_v4_64_128": {"params": nparams(cm.EnhancedPowerEncoder(64, 128, 2, 4)), "keys": keyshapes(cm.EnhancedPowerEncoder(64, 128, 2, 4))},
            "lite_erp_64_96": {"params": nparams(cm.LiteERPEncoder(64, 96)), "keys": keyshapes(cm.LiteERPEncoder(64, 96))},
            "lite_pw_64_96": {"params": nparams(cm.LitePowerEncoder(64, 96)), "keys": keyshapes(cm.LitePowerEncoder(64, 96))},
            "trimodal_lite_64_64_6048": {"params": nparams(cm.EnhancedTriModalFusionNetV4Lite(64, 64, 6048)),
                                         "keys": keyshapes(cm.EnhancedTriModalFusionNetV4Lite(64, 64, 6048))},
            "fmri_400_40000": {"params": nparams(fu.fMRIFusionNet(400, 40000)), "keys": keyshapes(fu.fMRIFusionNet(400, 40000))},
            "bridge_fusion_weights_init": bu.EEGfMRIBridgeFusionNet().get_fusion_weights(),
            "fmri_fusion_weights_init": fu.fMRIFusionNet(8, 8).get_fusion_weights(),
        }
    (OUT / "structure.json").write_text(json.dumps(struct, indent=1, sort_keys=True))
    print("wrote structure.json:", {k: v["params"] for k, v in struct.items() if isinstance(v, dict) and "params" in v})


if __name__ == "__main__":
    main()
