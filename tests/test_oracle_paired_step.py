"""CPU checks of the composite-step oracle (oracle/paired_step.py) and of the host-side bookkeeping of
PairedBridgeModel / synthetic data that the GPU parity tests rely on."""
import math

import torch

from multimodal_eeg_fmri_b200 import synthetic
from multimodal_eeg_fmri_b200.training import PairedBridgeModel
from oracle import infonce as oi
from oracle import paired_step as ps


def _model(encoder="v4"):
    torch.manual_seed(7)
    return PairedBridgeModel(eeg_channels=8, n_roi=12, eeg_hidden=32, fmri_hidden=16, bridge_dim=32, dropout=0.0,
                             fmri_dropout=0.0, encoder=encoder)


def test_contrastive_parameter_set_matches_oracle_trainable_keys():
    for enc in ("v4", "lite"):
        m = _model(enc)
        P = {k: v.clone() for k, v in m.state_dict().items()}
        ids = {id(p) for p in m.contrastive_parameters()}
        names = sorted(k for k, p in m.named_parameters() if id(p) in ids)
        assert names == sorted(ps.trainable_keys(P))
        assert len(ids) == len(names)


def test_oracle_paired_step_learns_on_a_fixed_batch():
    m = _model("lite")
    P = {k: v.clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(16, 8, 64, 12, 20, seed=5)
    state, losses = {}, []
    for _ in range(4):
        loss, total = ps.paired_train_step(P, state, eeg, roi, conn, 0.07, "lite")
        losses.append(float(loss))
        assert math.isfinite(float(total)) and float(total) > 0
    assert abs(losses[0] - math.log(16)) < 1.5
    assert losses[-1] < losses[0]


def test_sharded_loss_is_the_global_batch_loss():
    m = _model("lite")
    P = {k: v.clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(16, 8, 64, 12, 20, seed=5)
    shards = [(eeg[:8], roi[:8], conn[:8]), (eeg[8:], roi[8:], conn[8:])]
    a = ps.sharded_paired_loss(P, shards, 0.07, "lite")
    b = ps.paired_loss(P, eeg, roi, conn, 0.07, "lite")
    assert torch.equal(a, b)
    e, f = ps.paired_embeddings(P, eeg, roi, conn, "lite")
    tot, parts = oi.sharded_symmetric_infonce([e[:8], e[8:]], [f[:8], f[8:]], 0.07)
    assert abs(float(sum(parts)) - float(b)) < 1e-6 and abs(float(tot) - float(b)) < 1e-6


def test_synthetic_batches_are_deterministic_and_rank_offset():
    a = synthetic.paired_batch(4, 8, 32, 6, 10, seed=1)
    b = synthetic.paired_batch(4, 8, 32, 6, 10, seed=1)
    c = synthetic.paired_batch(4, 8, 32, 6, 10, seed=1, offset=4)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert not torch.equal(a[0], c[0])
    assert a[0].shape == (4, 8, 32) and a[1].shape == (4, 10, 6) and a[2].shape == (4, 36)
    r = synthetic.eeg_recordings(2, 4, 512)
    assert r.shape == (2, 4, 512) and r.dtype == torch.float32


def test_roi_connectivity_is_numpy_corrcoef():
    """Known-answer pin of the authored connectivity definition: numpy.corrcoef of the ROI columns, per sample,
    NaN -> 0 first; a constant column gives NaN in its row and column (numpy's 0 / 0)."""
    import numpy as np
    from oracle import models as om
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 40, 7, generator=g, dtype=torch.float64)
    x[1, 3, 2] = float("nan")
    got = om.roi_connectivity(x).reshape(3, 7, 7).numpy()
    for b in range(3):
        ref = np.corrcoef(np.nan_to_num(x[b].numpy()).T)
        assert np.allclose(got[b], ref, rtol=0, atol=1e-12)
    assert np.allclose(np.diagonal(got, axis1=1, axis2=2), 1.0)
    x[0, :, 4] = 2.5
    c = om.roi_connectivity(x).reshape(3, 7, 7)
    assert torch.isnan(c[0, 4]).all() and torch.isnan(c[0, :, 4]).all() and not torch.isnan(c[0, :4, :4]).any()


def test_paired_step_with_derived_connectivity_matches_explicit():
    """conn=None derives the connectivity from the ROI series: same loss as passing that matrix explicitly."""
    from oracle import models as om
    m = _model("lite")  # n_roi 12 -> conn_dim 144
    P = {k: v.clone() for k, v in m.state_dict().items()}
    eeg, roi, _ = synthetic.paired_batch(16, 8, 64, 12, 20, seed=5)
    a = ps.paired_loss(P, eeg, roi, None, 0.07, "lite")
    b = ps.paired_loss(P, eeg, roi, om.roi_connectivity(roi), 0.07, "lite")
    assert float(a) == float(b) and math.isfinite(float(a))
