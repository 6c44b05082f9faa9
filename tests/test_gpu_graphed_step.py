"""The whole training step as one CUDA-graph launch (recipe of _test_bridge.py:775-788):
`PairedTrainer.capture` / `GraphedStep`, the device-resident seed epoch behind it and the device-state AdamW.

Bit-exact comparisons: a replay runs the very kernels an eager step runs (the step is bit-reproducible,
tests/test_gpu_paired_step.py), so replay k must EQUAL the eager step made with the same host seeds at seed epoch k."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=3, dropout=0.3):
    """d_model 128 / 4 heads: the fused attention / FFN / residual-LayerNorm kernels of the BASELINE width, with dropout."""
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel
    torch.manual_seed(seed)
    return PairedBridgeModel(eeg_channels=8, n_roi=12, eeg_hidden=128, fmri_hidden=16, bridge_dim=32, dropout=dropout,
                             fmri_dropout=0.4 if dropout else 0.0, encoder="v4").cuda().train()


def _batch(B=32, seed=5):
    from multimodal_eeg_fmri_b200 import synthetic
    eeg, roi, _ = synthetic.paired_batch(B, 8, 64, 12, 20, seed=seed)
    return eeg.cuda(), roi.cuda()


@pytest.fixture(autouse=True)
def _epoch_zero():
    from multimodal_eeg_fmri_b200 import ops
    ops.seed_epoch_set(0)
    yield
    ops.seed_epoch_set(0)
    torch.cuda.synchronize()


def test_seed_epoch_selects_the_masks():
    """Epoch 0 leaves every mask as it was; another epoch draws other masks at the same keep rate; the epoch is read
    by every translation unit that hashes (elementwise, residual-LayerNorm, fused attention, fused FFN)."""
    from multimodal_eeg_fmri_b200 import ops
    assert ops.seed_epoch_get() == 0
    seed, p = 0x1234ABCD5678, 0.3
    x = torch.ones(512, 128, device="cuda")

    def masks():
        ffn = torch.empty(256, 512, device="cuda", dtype=torch.uint8)
        att = torch.empty(2 * 4 * 64 * 64, device="cuda", dtype=torch.uint8)
        ops._call("xm_ffn_fused_mask_u8", ops._p(ffn), 256, 512, p, seed, ops._stream())
        ops._call("xm_attn_fused_mask_u8", ops._p(att), 2, 64, 4, p, seed, ops._stream())
        ew = ops.act_fwd(x, "none", p, seed) != 0
        ln = ops.resid_seqmean_fwd(torch.zeros(4, 128, 128, device="cuda"), x.view(4, 128, 128).contiguous(), p, seed)
        return [ffn.clone(), att.clone(), ew, ln]  # ln: per (sample, column) mean of the kept, rescaled ones

    m0 = masks()
    ops.seed_epoch_set(7)
    assert ops.seed_epoch_get() == 7
    m7 = masks()
    ops.seed_epoch_advance()
    assert ops.seed_epoch_get() == 8
    m8 = masks()
    ops.seed_epoch_set(0)
    again = masks()
    for a, b, c, d in zip(m0, m7, m8, again):
        assert torch.equal(a, d)
        assert not torch.equal(a, b) and not torch.equal(b, c)
    for m in m7[:3]:  # element masks: keep rate 1 - p
        assert abs(float(m.float().mean()) - (1 - p)) < 0.02


def test_adamw_with_device_state_equals_host_state():
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(0)
    n = 100_003
    p0, g0 = torch.randn(n, device="cuda"), torch.randn(n, device="cuda") * 0.01
    outs = []
    for dev_state in (False, True):
        p, g, m, v = p0.clone(), g0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        step_dev = torch.zeros(1, device="cuda", dtype=torch.int64)
        lr_dev = torch.full((1,), 3e-4, device="cuda")
        for step in range(1, 4):
            g.copy_(g0 * step)
            if dev_state:
                step_dev.add_(1)
                norm = ops.clip_adamw_dev_(p, g, m, v, step_dev, lr_dev, 1e-2, 1.0)
            else:
                norm = ops.clip_adamw_(p, g, m, v, step, 3e-4, 1e-2, 1.0)
        outs.append((p, m, v, norm))
    for a, b in zip(*outs):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-9)


def test_capture_does_not_train_and_replays_equal_eager_steps():
    from multimodal_eeg_fmri_b200 import functional as XF, ops
    from multimodal_eeg_fmri_b200.training import PairedTrainer
    eeg, roi = _batch()
    eeg2, roi2 = _batch(seed=6)
    ma, mb = _model(), _model()
    ta, tb = PairedTrainer(ma, lr=1e-3), PairedTrainer(mb, lr=1e-3)
    tb.use_device_state()
    before = {k: v.clone() for k, v in ma.state_dict().items()}
    XF.manual_seed(99)
    g = ta.capture(eeg, roi)
    assert g.launches_captured > 100 and ta.step_count == 0 and ops.seed_epoch_get() == 0
    for k, v in ma.state_dict().items():
        assert torch.equal(v, before[k]), f"capture changed {k}"
    assert not ta.exp_avg.any() and not ta.exp_avg_sq.any()
    seeds = XF.seed_state()

    losses_g = [float(g(eeg, roi)), float(g(eeg2, roi2)), float(g(eeg, roi))]
    assert ops.seed_epoch_get() == 3 and ta.step_count == 3
    torch.cuda.synchronize()

    losses_e = []
    for k, (e, r) in enumerate([(eeg, roi), (eeg2, roi2), (eeg, roi)], start=1):
        XF.set_seed_state((seeds[0], 0))  # re-draw the host seeds the captured launches carry
        ops.seed_epoch_set(k)
        losses_e.append(float(tb.step(e, r)))
    assert losses_g == losses_e, (losses_g, losses_e)
    for (k, a), b in zip(ma.state_dict().items(), mb.state_dict().values()):
        assert torch.equal(a, b), f"{k} differs between replayed and eager steps"
    assert torch.equal(ta.exp_avg_sq, tb.exp_avg_sq)
    assert torch.equal(ta.last_grad_norm, tb.last_grad_norm)
    # fresh masks per replay: the same batch at epochs 1 and 3 gives different losses under dropout 0.3
    assert losses_g[0] != losses_g[2]


def test_replay_reads_learning_rate_and_step_count_from_the_device():
    from multimodal_eeg_fmri_b200.training import PairedTrainer
    eeg, roi = _batch()
    m = _model(dropout=0.0)
    tr = PairedTrainer(m, lr=1e-3)
    g = tr.capture(eeg, roi)
    l0 = float(g(eeg, roi))
    tr.lr = 0.0  # AdamW with lr 0 leaves the parameters (decay factor 1 - lr * wd = 1) but moves the moments
    frozen = tr.flat_param.clone()
    l1 = float(g(eeg, roi))
    assert torch.equal(tr.flat_param, frozen) and l1 < l0
    tr.lr = 1e-3
    l2 = float(g.replay())
    assert not torch.equal(tr.flat_param, frozen) and math.isclose(l2, l1, rel_tol=0, abs_tol=0)
    assert int(tr._step_dev) == tr.step_count == 3
    # an eager step after the replays continues the same state
    l3 = float(tr.step(eeg, roi))
    assert l3 < l2 and tr.step_count == 4 and int(tr._step_dev) == 4
    with pytest.raises(ValueError):
        g(eeg[:16], roi[:16])


def test_graphed_step_with_window_gather_and_host_inputs():
    """Raw recordings (window gather inside the graph) refilled from pinned host memory."""
    from multimodal_eeg_fmri_b200 import synthetic
    from multimodal_eeg_fmri_b200.training import PairedTrainer
    m = _model(dropout=0.1)
    tr = PairedTrainer(m, lr=1e-3, window=64, hop=64)
    rec = torch.randn(8, 8, 256).pin_memory()
    _, roi, _ = synthetic.paired_batch(32, 8, 64, 12, 20, seed=9)
    roi = roi.pin_memory()
    g = tr.capture(rec, roi)
    losses = [float(g(rec, roi)) for _ in range(6)]
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0]


def test_graphed_callable_replays_the_fmri_recipe():
    """run_fmri_v11.py:430-450's step (CE, clip_grad_norm_, torch AdamW) through GraphedCallable: constructing it does not
    train, replays equal eager steps at the same (host seeds, seed epoch), torch's capturable optimizer counts on the device."""
    from multimodal_eeg_fmri_b200 import functional as XF, ops
    from multimodal_eeg_fmri_b200.modules import fMRIFusionNet
    from multimodal_eeg_fmri_b200.training import GraphedCallable
    act, conn = torch.randn(64, 40, device="cuda"), torch.randn(64, 400, device="cuda")
    y = torch.randint(0, 2, (64,), device="cuda")

    def make():
        torch.manual_seed(1)
        m = fMRIFusionNet(40, 400).cuda().train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)

        def step(a, c, t):
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(m(a, c), t)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
            return loss.detach()
        return m, opt, step

    ma, oa, step_a = make()
    mb, ob, step_b = make()
    before = {k: v.clone() for k, v in ma.state_dict().items()}
    XF.manual_seed(5)
    g = GraphedCallable(step_a, [act, conn, y], [ma], [oa])
    for k, v in ma.state_dict().items():
        assert torch.equal(v, before[k]), k
    base = XF.seed_state()[0]
    lg = [float(g(act, conn, y)) for _ in range(3)]
    le = []
    for k in (1, 2, 3):
        XF.set_seed_state((base, 0))
        ops.seed_epoch_set(k)
        le.append(float(step_b(act, conn, y)))
    assert lg == le, (lg, le)
    for (k, a), b in zip(ma.state_dict().items(), mb.state_dict().values()):
        assert torch.equal(a, b), k
    assert all(math.isfinite(v) for v in lg) and len(set(lg)) == 3  # three different steps (parameters and masks move)


def test_lite_wrapper_is_capturable_and_its_fusion_weights_follow_the_replays():
    """run_training_lite.py:302-328's wrapper asks for the fusion weights in every forward; they stay on the device
    (modules.DeviceFloats), so the step (run_training_lite.py:478-489) can be captured, and get_fusion_weights() read after a
    replay equals what the eager twin reports after the same step."""
    from multimodal_eeg_fmri_b200 import functional as XF, ops
    from multimodal_eeg_fmri_b200.modules import LabelSmoothingCrossEntropy
    from multimodal_eeg_fmri_b200.run_training_lite import ImprovedTriModalFusionNetLite
    from multimodal_eeg_fmri_b200.training import GraphedCallable
    erp, pw, cn = torch.randn(16, 8, 64, device="cuda"), torch.randn(16, 6, 64, device="cuda"), torch.randn(16, 40, device="cuda")
    y = torch.randint(0, 2, (16,), device="cuda")
    crit = LabelSmoothingCrossEntropy(0.1)

    def make():
        torch.manual_seed(2)
        m = ImprovedTriModalFusionNetLite(6, 8, 40, fusion_dim=32).cuda().train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.01, capturable=True)

        def step(p, e, c, t):
            opt.zero_grad(set_to_none=True)
            loss = crit(m(p, e, c), t)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
            return loss.detach()
        return m, opt, step

    ma, oa, step_a = make()
    mb, ob, step_b = make()
    XF.manual_seed(6)
    g = GraphedCallable(step_a, [pw, erp, cn, y], [ma], [oa])
    base = XF.seed_state()[0]
    for k in (1, 2):
        lg = float(g(pw, erp, cn, y))
        wg = dict(ma.get_fusion_weights())
        XF.set_seed_state((base, 0))
        ops.seed_epoch_set(k)
        le = float(step_b(pw, erp, cn, y))
        we = dict(mb.get_fusion_weights())
        ops.seed_epoch_set(k)
        assert lg == le and wg == we and set(wg) == {"erp_weight", "pw_weight", "conn_weight"}, (k, lg, le, wg, we)
    for (k, a), b in zip(ma.state_dict().items(), mb.state_dict().values()):
        assert torch.equal(a, b), k
