"""Host-side logic above the C ABI, on the CPU with the TEST-ONLY fake device ops (tests/fake_ops.py):
autograd functions + drop-in modules against the oracle, and the data-parallel step (SyncBN partial
sums, embedding all-gather with exact local gradients, one flat gradient bucket) under a
world_size-2 gloo group against the single-process global-batch run."""
import socket

import pytest
import torch
import torch.multiprocessing as mp

import dp_worker
import fake_ops
from conftest import assert_close_rel
from multimodal_eeg_fmri_b200 import synthetic
from oracle import paired_step as ps


@pytest.fixture()
def fakes(monkeypatch):
    fake_ops.install(monkeypatch)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("encoder", ["v4", "lite"])
def test_paired_model_autograd_plumbing_vs_oracle(fakes, encoder):
    m = dp_worker.make_model(encoder).train()
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(16, 8, 64, 12, 20, seed=3)
    loss = m(eeg, roi, conn)
    loss.backward()
    oloss, ograds = ps.paired_loss_and_grads(P, eeg, roi, conn, 0.07, encoder)
    assert_close_rel(loss, oloss, 1e-5, "loss")
    named = dict(m.named_parameters())
    for k, g in ograds.items():
        assert_close_rel(named[k].grad, g, 2e-4, f"grad {k}", atol=1e-6)


def test_paired_model_with_derived_connectivity_vs_oracle(fakes):
    """conn=None: the connectivity features are the flattened ROI x ROI correlation matrix of the series, derived
    inside the step (n_roi 12 -> conn_dim 144), on both sides."""
    m = dp_worker.make_model("lite").train()
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    eeg, roi, _ = synthetic.paired_batch(16, 8, 64, 12, 20, seed=3)
    loss = m(eeg, roi)
    loss.backward()
    oloss, ograds = ps.paired_loss_and_grads(P, eeg, roi, None, 0.07, "lite")
    assert_close_rel(loss, oloss, 1e-5, "loss")
    named = dict(m.named_parameters())
    for k, g in ograds.items():
        assert_close_rel(named[k].grad, g, 2e-4, f"grad {k}", atol=1e-6)


def test_trimodal_lite_and_fmri_modules_vs_golden(fakes):
    from conftest import load_golden
    from multimodal_eeg_fmri_b200 import modules
    g = load_golden("trimodal_lite_small")
    m = modules.EnhancedTriModalFusionNetV4Lite(8, 8, 30, hidden_dim=24, num_classes=2, dropout=0.0, conn_boost=1.3)
    m.load_state_dict(g["sd"], strict=True)
    m.train()
    logits, w, fused = m(g["inputs"][0], g["inputs"][1], g["inputs"][2], return_fusion_weights=True, return_fused_feats=True)
    assert_close_rel(logits, g["outputs"][0], 1e-5, "logits")
    assert_close_rel(fused, g["raw"]["fused"], 1e-5, "fused")
    assert_close_rel(torch.tensor([w["erp_weight"], w["pw_weight"], w["conn_weight"]]), g["raw"]["weights"], 1e-5, "weights")
    g = load_golden("bridge_small")
    b = modules.EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    b.load_state_dict(g["sd"], strict=True)
    b.eval()
    logits, fused, fw, aw = b(g["inputs"][0], g["inputs"][1], return_features=True, return_weights=True)
    assert_close_rel(logits, g["outputs"][0], 1e-5, "bridge logits")
    assert_close_rel(aw, g["raw"]["attn_weights"], 1e-5, "attention weights")
    assert_close_rel(fw, g["raw"]["fusion_weights"], 1e-5, "fusion weights")


def test_trainer_matches_oracle_recipe(fakes):
    from multimodal_eeg_fmri_b200.training import PairedTrainer
    m = dp_worker.make_model("lite").train()
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(16, 8, 64, 12, 20, seed=5)
    tr = PairedTrainer(m)
    state = {}
    for _ in range(3):
        a = float(tr.step(eeg, roi, conn))
        b = float(ps.paired_train_step(P, state, eeg, roi, conn, 0.07, "lite")[0])
        assert abs(a - b) < 1e-5 * abs(b)
    for k in set(ps.trainable_keys(P)) - set(ps.bias_before_batchnorm_keys(P)):
        assert_close_rel(m.state_dict()[k], P[k], 1e-5, k, atol=2e-5)


@pytest.mark.parametrize("encoder", ["lite", "v4"])
def test_two_rank_gloo_step_equals_global_batch_step(tmp_path, encoder):
    """Rank r trains on rows [8r, 8r+8) of a 16-sample batch; the run must reproduce the
    single-process 16-sample run: same global loss (sum of shares) and same parameters."""
    steps = 2
    mp.spawn(dp_worker.run, args=(2, _free_port(), encoder, steps, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(tmp_path / "dp.pt")
    m = dp_worker.make_model(encoder)
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(16, 8, 64, 12, 20, seed=5)
    state, want = {}, []
    for _ in range(steps):
        want.append(float(ps.paired_train_step(P, state, eeg, roi, conn, 0.07, encoder)[0]))
    assert_close_rel(torch.tensor(got["losses"]), torch.tensor(want), 1e-5, "global loss per step")
    for k in set(ps.trainable_keys(P)) - set(ps.bias_before_batchnorm_keys(P)):
        assert_close_rel(got["sd"][k], P[k], 1e-5, f"param {k}", atol=2e-5)


def test_fused_transformer_tail_autograd_vs_oracle(fakes):
    """The hand-written backward of the fused PE + blocks + mean-pool tail (XF.TransformerTail) at a shape the
    fused kernels support (d_model 128, 4 heads of 32, L = 25): outputs and every gradient vs the oracle."""
    from multimodal_eeg_fmri_b200 import modules
    from oracle import models as om
    torch.manual_seed(3)
    m = modules.EnhancedERPEncoder(8, 128, 2, 4, 0.0).train()
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(6, 8, 50)
    cot = torch.randn(6, 128)
    xg = x.clone().requires_grad_(True)
    y = m(xg)
    y.backward(cot)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")}
    xo = x.clone().requires_grad_(True)
    yo = om.enhanced_erp_encoder({**P, **leaves}, "", xo, nhead=4)
    yo.backward(cot)
    assert_close_rel(y, yo, 1e-5, "encoder output")
    assert_close_rel(xg.grad, xo.grad, 2e-4, "dx")
    for k, p in m.named_parameters():
        assert_close_rel(p.grad, leaves[k].grad, 2e-4, f"grad {k}", atol=1e-6)


def _xai_fixture():
    import numpy as np
    from conftest import GOLDEN
    from multimodal_eeg_fmri_b200 import bridge_utils as bu
    z = np.load(GOLDEN / "bridge_xai.npz")
    m = bu.EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    m.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}, strict=True)
    return bu, z, m, torch.from_numpy(z["eeg"]), torch.from_numpy(z["fmri"]), torch.from_numpy(z["given_target"])


def check_bridge_xai(device, tol):
    """Shared by the CPU (fake ops) and GPU tests: batched saliency / integrated gradients / weight extraction
    against the REAL reference classes' outputs (tests/golden/bridge_xai.npz, bridge_utils.py:158-270)."""
    import numpy as np
    bu, z, m, eeg, fmri, given = _xai_fixture()
    m = m.to(device)
    sal = bu.BridgeGradientSaliency(m, device)
    for tag, tc in (("pred", None), ("given", given)):
        r = sal.compute(eeg, fmri, tc)
        assert_close_rel(r["eeg"], z[f"saliency_{tag}/eeg"], tol, f"saliency {tag} eeg")
        assert_close_rel(r["fmri"], z[f"saliency_{tag}/fmri"], tol, f"saliency {tag} fmri")
    for n in (50, 7):
        ig = bu.BridgeIntegratedGradients(m, device, n_steps=n)
        for tag, tc in (("pred", None), ("given", given)):
            r = ig.compute(eeg, fmri, tc)
            assert r["eeg"].shape == (6, 32) and r["fmri"].shape == (6, 16)
            assert_close_rel(r["eeg"], z[f"ig{n}_{tag}/eeg"], tol, f"IG{n} {tag} eeg")
            assert_close_rel(r["fmri"], z[f"ig{n}_{tag}/fmri"], tol, f"IG{n} {tag} fmri")
    labels = {s: int(s % 2) for s in range(1, 7)}
    ds = bu.BridgeFeatureDataset({s: eeg[s - 1] for s in labels}, {str(s): fmri[s - 1] for s in labels}, labels,
                                 [3, 1, 2, 6, 5, 4, 9])
    rows = bu.extract_attention_and_fusion_weights(m, ds, device)
    assert [r["subject"] for r in rows] == z["extract/subject"].tolist()  # bit-exact integer work
    assert [r["label"] for r in rows] == z["extract/label"].tolist()
    assert [r["prediction"] for r in rows] == z["extract/prediction"].tolist()
    assert rows[0]["fusion_weights"].shape == (2,) and rows[0]["attn_weights"].shape == (2,)
    assert_close_rel(np.stack([r["fusion_weights"] for r in rows]), z["extract/fusion_weights"], tol, "fusion weights")
    assert_close_rel(np.stack([r["attn_weights"] for r in rows]), z["extract/attn_weights"], tol, "attention weights")
    assert bu.extract_attention_and_fusion_weights(m, bu.BridgeFeatureDataset({}, {}, {}, []), device) == []


def test_bridge_attribution_helpers_vs_reference_golden(fakes):
    check_bridge_xai("cpu", 1e-5)


def test_shard_batches_feed_the_trainer_identically(fakes, tmp_path):
    """Formats -> step: batches read back from an XMSHARD1 file drive the trainer to bit-identical losses as the
    in-memory tensors they were written from (and two ranks' shares interleave to the full batch sequence)."""
    from multimodal_eeg_fmri_b200 import shards
    from multimodal_eeg_fmri_b200.training import PairedTrainer
    eeg, roi, conn = synthetic.paired_batch(48, 8, 64, 12, 20, seed=11)
    shards.write_shard(tmp_path / "paired.xms", {"eeg": eeg, "roi": roi, "conn": conn}, meta={"seed": 11})
    sh = shards.Shard(tmp_path / "paired.xms")
    names = ["eeg", "roi", "conn"]
    ta, tb = PairedTrainer(dp_worker.make_model("lite").train()), PairedTrainer(dp_worker.make_model("lite").train())
    from_file = [float(ta.step(*b)) for b in shards.host_batches(sh, names, 16, pin=False)]
    in_memory = [float(tb.step(eeg[i:i + 16], roi[i:i + 16], conn[i:i + 16])) for i in range(0, 48, 16)]
    assert from_file == in_memory and len(from_file) == 3
    r0 = list(shards.host_batches(sh, names, 16, pin=False, ranks=(0, 2)))
    r1 = list(shards.host_batches(sh, names, 16, pin=False, ranks=(1, 2)))
    # 3 batches over 2 ranks: both ranks get ONE batch (an uneven split would leave rank 0 alone in its step's collectives)
    assert len(r0) == len(r1) == 1 and torch.equal(r0[0][0], eeg[:16]) and torch.equal(r1[0][0], eeg[16:32])
