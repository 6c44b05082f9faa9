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


def case_module(name, module, inputs, extra=None, seed=0, train=True, skip=()):
    """`skip`: state_dict key suffixes left out of the fixture (deterministic buffers such as the 2.5 MB sinusoidal
    table `pos_encoder.pe`, which the module under test rebuilds identically)."""
    store = {}
    sd0 = {k: v.clone() for k, v in module.state_dict().items() if not k.endswith(tuple(skip) or ("\0",))}
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


def case_bridge_xai(bu):
    """Attribution helpers of the reference (bridge_utils.py:158-270) on the `bridge_small` model: gradient
    saliency (predicted and given target), integrated gradients (50 and 7 steps; the default target is fixed by
    the alpha = 0 pass) and the per-subject attention / fusion-weight extraction."""
    torch.manual_seed(48)
    br = bu.EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    g = torch.Generator().manual_seed(4848)
    eeg, fmri = torch.randn(6, 32, generator=g), torch.randn(6, 16, generator=g)
    given = torch.tensor([1, 0, 1, 1, 0, 0])
    store = {}
    _pack(store, "sd", br.state_dict())
    store.update({"eeg": _np(eeg), "fmri": _np(fmri), "given_target": given.numpy()})
    sal = bu.BridgeGradientSaliency(br, "cpu")
    for tag, tc in (("pred", None), ("given", given)):
        r = sal.compute(eeg, fmri, tc)
        store[f"saliency_{tag}/eeg"], store[f"saliency_{tag}/fmri"] = r["eeg"], r["fmri"]
    for n in (50, 7):
        ig = bu.BridgeIntegratedGradients(br, "cpu", n_steps=n)
        for tag, tc in (("pred", None), ("given", given)):
            r = ig.compute(eeg, fmri, tc)
            store[f"ig{n}_{tag}/eeg"], store[f"ig{n}_{tag}/fmri"] = r["eeg"], r["fmri"]
    labels = {s: int(s % 2) for s in range(1, 7)}
    ds = bu.BridgeFeatureDataset({s: eeg[s - 1] for s in labels}, {str(s): fmri[s - 1] for s in labels}, labels, [3, 1, 2, 6, 5, 4, 9])
    rows = bu.extract_attention_and_fusion_weights(br, ds, "cpu")
    store["extract/subject"] = np.array([r["subject"] for r in rows])
    store["extract/label"] = np.array([r["label"] for r in rows])
    store["extract/prediction"] = np.array([r["prediction"] for r in rows])
    store["extract/fusion_weights"] = np.stack([r["fusion_weights"] for r in rows])
    store["extract/attn_weights"] = np.stack([r["attn_weights"] for r in rows])
    np.savez_compressed(OUT / "bridge_xai.npz", **store)
    print(f"wrote bridge_xai.npz ({len(store)} arrays)")


def _import_reference_eeg_data_utils():
    """EEG_CODE/eeg_data_utils.py imports h5py, which this image lacks.  The stub's File raises OSError, which is
    what the real h5py does for a MATLAB v5 (non-HDF5) file -- the only kind the fixtures contain -- so the
    reference takes its own `scipy.io.loadmat` fallback branch (eeg_data_utils.py:170-183)."""
    if "h5py" not in sys.modules:
        stub = types.ModuleType("h5py")

        def _not_hdf5(*a, **k):
            raise OSError("Unable to open file (file signature not found)")

        stub.File = _not_hdf5
        sys.modules["h5py"] = stub
    from EEG_CODE import eeg_data_utils as edu
    return edu


def _reference_bridge_raw_dataset():
    """`_test_bridge.py` runs the whole pipeline at import, so only the source lines of its BridgeRawDataset class
    (:391-453) are executed, in a namespace holding the names the class uses."""
    import logging
    from collections import defaultdict
    from torch.utils.data import Dataset
    lines = (REF / "_test_bridge.py").read_text().splitlines()
    a = next(i for i, l in enumerate(lines) if l.startswith("class BridgeRawDataset(Dataset):"))
    b = next(i for i, l in enumerate(lines) if l.startswith("bridge_raw_dataset = BridgeRawDataset("))
    ns = {"Dataset": Dataset, "defaultdict": defaultdict, "np": np, "logger": logging.getLogger("ref")}
    exec("\n".join(lines[a:b]), ns)
    return ns["BridgeRawDataset"]


def case_loaders(fu):
    """On-disk formats either side of the path: writes a small fixture tree under
    tests/golden/data/ (fMRI CSVs, label CSVs, MATLAB v5 files) and stores what the REFERENCE loaders
    (fmri_utils.py:115-241, eeg_data_utils.py:19-186) and BridgeRawDataset (_test_bridge.py:391-453) return."""
    import shutil
    from scipy.io import savemat
    root = OUT / "data"
    if root.exists():
        shutil.rmtree(root)
    rng = np.random.default_rng(20261018)
    fm = root / "fmri"
    subjects = [1, 2, 3, 5]  # 4 has no directory; 5 has only a broken file
    for s in (1, 2, 3):
        d = fm / f"sub-{s}"
        d.mkdir(parents=True)
        for t, (tr, roi) in (("taskA", (6, 4)), ("taskB", (5, 3))):
            if s == 3 and t == "taskB":
                continue  # missing type for one subject
            x = rng.standard_normal((tr, roi)).round(4)
            cols = {f"roi{j}": x[:, j] for j in range(roi)}
            lines_ = [",".join((["Subject"] if s == 2 else []) + list(cols))]
            for r in range(tr):
                cells = [("" if (s == 1 and r == 2 and j == 1) else repr(float(x[r, j]))) for j in range(roi)]  # one NaN cell
                lines_.append(",".join(([str(s)] if s == 2 else []) + cells))
            (d / f"subject_{s}_activation_{t}.csv").write_text("\n".join(lines_) + "\n")
        for t in ("rest", "task"):
            if s == 2 and t == "task":
                continue
            m = rng.standard_normal((4, 4)).round(4)
            rows = [",".join(f"r{j}" for j in range(4))] + [",".join("" if (i == j == 3) else repr(float(m[i, j])) for j in range(4)) for i in range(4)]
            (d / f"subject_{s}_fdr_PPI_Connectivity_{t}.csv").write_text("\n".join(rows) + "\n")
    d5 = fm / "sub-5"
    d5.mkdir()
    (d5 / "subject_5_activation_taskA.csv").write_text("a,b\nx,y\nz,w\n")  # not numeric: logged and skipped
    lab1, lab2, lab3 = root / "labels_str", root / "labels_num" / "inner", root / "labels_bad"
    for d in (lab1, lab2, lab3):
        d.mkdir(parents=True)
    (lab1 / "labels.csv").write_text("Subject,Outcome\n1,Good\n2,bad\n3,YES\n7,positive\n5,1\n")
    (lab2.parent / "labels.csv").write_text("id,group\n1,0\n2,1\n3,3\n9,1\n")  # found through label_path.parent
    (lab3 / "outcomes.csv").write_text("who,what\n1,2\n")
    eeg = root / "eeg"
    for sub in ("conn", "pw", "erp", "labels"):
        (eeg / sub).mkdir(parents=True)
    def mat(path, arr):
        savemat(path, {"data": arr})
    nan = np.float32("nan")
    mat(eeg / "conn" / "conn_Alpha_open_sub01.mat", np.array([[1.0, nan], [0.5, 2.0]], dtype=np.float32))
    mat(eeg / "conn" / "conn_alpha_close_sub01.mat", rng.standard_normal((2, 2)).astype(np.float32))  # band_key fallback name
    mat(eeg / "conn" / "conn_Beta_open_sub02.mat", rng.standard_normal((2, 2)).astype(np.float32))
    mat(eeg / "conn" / "conn_Alpha_open_sub03.mat", rng.standard_normal((2, 2)).astype(np.float32))
    for s_, band, freq in ((1, "alpha", "1_Hz"), (1, "alpha", "2_Hz"), (2, "beta", "1_Hz")):
        mat(eeg / "pw" / f"powspctrm_{band}_{freq}_sub{s_:02d}.mat", rng.standard_normal((3, 2)).astype(np.float64))
    for s_, band, freq, suffix in ((1, "alpha", "1_Hz", ""), (1, "alpha", "2_Hz", "_v2"), (2, "beta", "1_Hz", ""), (3, "alpha", "1_Hz", "")):
        a = rng.standard_normal((3, 5)).astype(np.float32)
        a[0, 0] = nan
        mat(eeg / "erp" / f"ERP_sub{s_:02d}_{band}_{freq}{suffix}.mat", a)
    (eeg / "erp" / "ERP_sub02_beta_2_Hz.mat").write_bytes(b"not a mat file")  # both readers fail: logged
    (eeg / "labels" / "medical_score.csv").write_text("Subject,Postoperative evaluation\nsub01,1\nsub02,3\nsub03,\nsub05,2\nsub07,4\n")

    edu = _import_reference_eeg_data_utils()
    store = {}
    def put(prefix, d):
        for k, v in d.items():
            key = k if not isinstance(k, tuple) else "|".join(map(str, k))
            store[f"{prefix}/{key}"] = np.asarray(v)
    with contextlib.redirect_stderr(io.StringIO()):  # tqdm bars
        for agg in ("both", "mean", "std", "median"):
            put(f"act_{agg}", fu.load_activation_features(fm, subjects, ["taskA", "taskB"], agg))
        put("act_both_B_only", fu.load_activation_features(fm, subjects, ["taskB"], "both"))
        put("conn", fu.load_connectivity_features(fm, subjects, ["rest", "task"]))
    put("labels_str", fu.load_fmri_labels(lab1, [1, 2, 3, 5]))
    put("labels_num", fu.load_fmri_labels(lab2, [1, 2, 3]))
    try:
        fu.load_fmri_labels(lab3, [1])
        raise AssertionError("expected ValueError")
    except ValueError as e:
        store["labels_bad_error"] = np.array(str(e).replace(str(lab3), "<dir>"))
    bands = {"alpha": "Alpha", "beta": "Beta"}
    e_conn = edu.load_eeg_conn_features(eeg / "conn", [1, 2, 3], bands, ["open", "close"])
    e_pw = edu.load_eeg_pw_features(eeg / "pw", [1, 2, 3], ["alpha", "beta"], ["1_Hz", "2_Hz"])
    e_erp = edu.load_eeg_erp_features(eeg / "erp", [1, 2, 3], ["alpha", "beta"], ["1_Hz", "2_Hz"])
    put("eeg_conn", e_conn); put("eeg_pw", e_pw); put("eeg_erp", e_erp)
    # the reference tests `dtype == object` for 'subNN' ids (eeg_data_utils.py:34): pandas >= 3 infers `str`
    # instead and the reference then fails on 'sub01'; run it with the inference it was written for
    import pandas as pd
    with pd.option_context("future.infer_string", False):
        put("eeg_labels_binary", edu.load_eeg_labels(eeg / "labels", True))
        put("eeg_labels_raw", edu.load_eeg_labels(eeg / "labels", False))
        labels = edu.load_eeg_labels(eeg / "labels", True)
    # BridgeRawDataset on those dicts
    RawDS = _reference_bridge_raw_dataset()
    with contextlib.redirect_stderr(io.StringIO()):
        f_act = fu.load_activation_features(fm, subjects, ["taskA", "taskB"], "both")
        f_conn = fu.load_connectivity_features(fm, subjects, ["rest", "task"])
    labels[3] = 1
    ds = RawDS(e_erp, e_pw, e_conn, f_act, f_conn, labels, ["3", "2", "1", "5", "4", "10"], bands, ["open", "close"])
    store["raw_ds/subjects"] = np.array([ds[i][4] for i in range(len(ds))])
    store["raw_ds/labels"] = np.array([ds[i][3] for i in range(len(ds))])
    store["raw_ds/n_eeg"] = np.array([len(ds[i][0]) for i in range(len(ds))])
    for i in range(len(ds)):
        for j, (erp, pw, conn) in enumerate(ds[i][0]):
            store[f"raw_ds/{i}/{j}/erp"], store[f"raw_ds/{i}/{j}/pw"], store[f"raw_ds/{i}/{j}/conn"] = erp, pw, conn
        store[f"raw_ds/{i}/fmri_act"], store[f"raw_ds/{i}/fmri_conn"] = _np(ds[i][1]), _np(ds[i][2])
    np.savez_compressed(OUT / "loaders.npz", **store)
    print(f"wrote loaders.npz ({len(store)} arrays) and the fixture tree {root}")


def main():
    cm, em, fu, bu = import_reference()
    only = [a for a in sys.argv[1:] if a in ("xai", "loaders", "d128")]  # regenerate only these fixtures
    if only:
        if "xai" in only:
            case_bridge_xai(bu)
        if "loaders" in only:
            case_loaders(fu)
        if "d128" in only:
            # the v4 ERP encoder at the BASELINE width (d_model 128, 4 heads of 32, 2 blocks): the shape at which the
            # CUDA path runs its fused transformer tail (fused attention core, fused FFN, fused residual + LayerNorm)
            torch.manual_seed(52)
            g = torch.Generator().manual_seed(52)
            erp = em.EnhancedERPEncoder(16, 128, 2, 4, 0.0)
            x = torch.randn(4, 16, 96, generator=g)
            erp.train()
            case_module("erp_v4_d128", erp, [x], extra={"conv_stack_out": _np(erp.conv_layers(x))}, seed=11,
                        skip=("pos_encoder.pe",))
        return
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

    case_bridge_xai(bu)
    case_loaders(fu)

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
