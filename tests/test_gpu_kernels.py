"""Per-kernel parity on a B200, called through the C ABI (multimodal_eeg_fmri_b200.ops binds
include/xmodal_b200.h with ctypes).  References are fp64 torch restatements of the same op.

Tolerances (BASELINE.json north_star): tf32 tensor-core contractions 1e-3 relative; fp32
reductions / norms / spectral 1e-5 relative; integer index work bit-exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close_rel, rel_err

pytestmark = pytest.mark.gpu

TF32 = 1e-3
FP32 = 1e-5


@pytest.fixture(scope="module")
def ops():
    from multimodal_eeg_fmri_b200 import _lib, ops as _ops
    assert _lib.LIB_PATH.exists(), "CUDA library missing on the GPU box"
    return _ops


def _nwc(x):
    return x.transpose(1, 2).contiguous()


# ------------------------------------------------------------------ linear
@pytest.mark.parametrize("M,N,K,act,splits", [
    (128, 64, 64, None, 1), (300, 96, 200, "gelu", 1), (4096, 128, 400, "relu", 1), (256, 128, 4096, None, 4),
    (70, 2, 32, None, 1), (512, 256, 128, None, 1), (1000, 320, 96, None, 1), (64, 128, 40000, None, 0),
    (1, 16, 8, None, 1),
])
def test_linear_fwd(ops, M, N, K, act, splits):
    torch.manual_seed(0)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    y = ops.linear_fwd(x, w, b, act=act, splits=splits)
    ref = x.double() @ w.double().t() + b.double()
    ref = F.gelu(ref) if act == "gelu" else torch.relu(ref) if act == "relu" else ref
    assert_close_rel(y, ref, TF32, "linear_fwd")


@pytest.mark.parametrize("M,N,K", [(128, 32, 32), (300, 96, 200), (4096, 128, 400), (256, 64, 128), (512, 2, 64),
                                   (64, 128, 40000)])
def test_linear_dgrad(ops, M, N, K):
    torch.manual_seed(1)
    dy = torch.randn(M, N, device="cuda")
    w = torch.randn(N, K, device="cuda")
    assert_close_rel(ops.linear_dgrad(dy, w), dy.double() @ w.double(), TF32, "linear_dgrad")


@pytest.mark.parametrize("M,N,K,splits", [(128, 128, 32, 1), (300, 96, 200, 1), (4096, 128, 400, 4), (256, 64, 128, 1),
                                          (2048, 128, 4000, 2), (512, 2, 64, 1), (64, 128, 40000, 0)])
def test_linear_wgrad(ops, M, N, K, splits):
    torch.manual_seed(2)
    dy = torch.randn(M, N, device="cuda")
    x = torch.randn(M, K, device="cuda")
    dw, db = ops.linear_wgrad(dy, x, splits=splits)
    assert_close_rel(dw, dy.double().t() @ x.double(), TF32, "linear_wgrad")
    assert_close_rel(db, dy.double().sum(0), FP32, "bias grad")


@pytest.mark.parametrize("M,N,K,act", [(300, 96, 200, "relu"), (2048, 128, 40000, None), (130, 64, 4000, None),
                                       (256, 16, 1000, None), (64, 256, 3000, "gelu")])
def test_linear_precise_is_fp32_accurate(ops, M, N, K, act):
    """3-pass tf32 + chunked fp32 accumulation: the projections whose outputs feed ReLU masks / BatchNorm
    statistics must be fp32-accurate at any K (a one-sided error of 4e-5 flips ~20 ReLU masks of the 40 000-d
    connectivity layer at batch 2048 and shows up as a 1e-2 error of that layer's gradients)."""
    torch.manual_seed(5)
    x = torch.randn(M, K, device="cuda") + 0.3  # a common component: the accumulators keep their sign
    w = torch.randn(N, K, device="cuda") / K ** 0.5 + 0.2 / K ** 0.5
    b = torch.randn(N, device="cuda")
    y = ops.linear_fwd_precise(x, w, b, act=act)
    ref = x.double() @ w.double().t() + b.double()
    ref = F.gelu(ref) if act == "gelu" else torch.relu(ref) if act == "relu" else ref
    assert_close_rel(y, ref, 3e-6, "linear_fwd_precise")
    dy = torch.randn(M, N, device="cuda")
    assert_close_rel(ops.linear_dgrad_precise(dy, w), dy.double() @ w.double(), 3e-6, "linear_dgrad_precise")
    dw, db = ops.linear_wgrad_precise(dy, x)
    assert_close_rel(dw, dy.double().t() @ x.double(), 3e-5, "linear_wgrad_precise")
    assert_close_rel(db, dy.double().sum(0), FP32, "bias grad")
    # the shared row-stacked split: same forward, same weight gradient
    x3 = ops.linear_precise_prepare(x)
    assert_close_rel(ops.linear_fwd_prepared(x3, w, b, act=act), ref, 3e-6, "linear_fwd_prepared")
    dw2, _ = ops.linear_wgrad_prepared(dy, x3)
    assert torch.equal(dw2, dw)


@pytest.mark.parametrize("Ml,Ng,D", [(256, 256, 128), (512, 4096, 128), (100, 36, 64), (1, 1, 128), (130, 8200, 32)])
def test_infonce_dgrad_is_fp32_accurate(ops, Ml, Ng, D):
    """dx = G @ f_n with G = softmax - onehot (rows sum to ~0, so the product cancels against the common
    component of the unit vectors)."""
    torch.manual_seed(6)
    f = F.normalize(torch.randn(Ng, D, device="cuda") + 0.7, dim=1)
    G = torch.softmax(torch.randn(Ml, Ng, device="cuda") * 3, dim=1)
    G[torch.arange(Ml), torch.arange(Ml) % Ng] -= 1.0
    for which in (0, 1):
        fn, f3, _ = ops.l2norm_split_fwd(f, which)
        dx = ops.infonce_dgrad(G, f3, which)
        assert_close_rel(dx, G.double() @ fn.double(), 5e-6, "infonce_dgrad", atol=1e-7)


# ------------------------------------------------------------------ conv1d
CONV_CASES = [(1, 32, 32, 128, 1), (2, 32, 32, 128, 3), (3, 64, 64, 500, 7), (2, 64, 128, 500, 5), (2, 128, 128, 250, 3),
              (2, 48, 96, 250, 5), (2, 18, 48, 500, 7), (1, 192, 128, 100, 1), (2, 8, 16, 64, 7), (5, 64, 64, 37, 5)]


@pytest.mark.parametrize("B,Cin,Cout,T,k", CONV_CASES)
def test_conv1d_fwd_dgrad_wgrad(ops, B, Cin, Cout, T, k):
    torch.manual_seed(3)
    x = torch.randn(B, Cin, T, device="cuda")
    w = torch.randn(Cout, Cin, k, device="cuda") / (Cin * k) ** 0.5
    b = torch.randn(Cout, device="cuda")
    dy = torch.randn(B, Cout, T, device="cuda")
    wk, wt = ops.conv1d_pack_weight(w)
    xl = ops.to_nwc(x)
    assert torch.equal(xl, _nwc(x)), "layout change must be exact"
    y = ops.conv1d_fwd(xl, wk, b, Cout)
    assert_close_rel(y, _nwc(F.conv1d(x.double(), w.double(), b.double(), padding=k // 2)), TF32, "conv fwd")
    dx = ops.conv1d_dgrad(_nwc(dy), wt, Cin)
    assert_close_rel(dx, _nwc(F.conv_transpose1d(dy.double(), w.double(), padding=k // 2)), TF32, "conv dgrad")
    dw, db = ops.conv1d_wgrad(_nwc(dy), xl, k)
    wd = torch.zeros(Cout, Cin, k, device="cuda", dtype=torch.float64, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv1d(x.double(), wd, None, padding=k // 2), wd, dy.double())
    assert_close_rel(dw, gw, TF32, "conv wgrad")
    assert_close_rel(db, dy.double().sum((0, 2)), FP32, "conv bias grad", atol=1e-5)


@pytest.mark.parametrize("B,Cin,Cout,T,k", [(3, 64, 64, 500, 7), (2, 64, 128, 250, 5), (300, 64, 64, 500, 7), (2, 18, 48, 37, 3),
                                            (5, 192, 128, 100, 1)])
def test_conv1d_fwd_emits_batchnorm_statistics(ops, B, Cin, Cout, T, k):
    """The statistics epilogue: same y, and partial sums whose total is (sum y, sum y^2) per channel over (B, T)."""
    torch.manual_seed(4)
    x = _nwc(torch.randn(B, Cin, T, device="cuda"))
    w = torch.randn(Cout, Cin, k, device="cuda") / (Cin * k) ** 0.5
    b = torch.randn(Cout, device="cuda")
    wk, _ = ops.conv1d_pack_weight(w)
    y0 = ops.conv1d_fwd(x, wk, b, Cout)
    y, part = ops.conv1d_fwd(x, wk, b, Cout, stats=True)
    assert torch.equal(y, y0)
    tot = part.sum(0)  # (Cout, 2) fp64
    yd = y.double().reshape(-1, Cout)
    assert_close_rel(tot[:, 0], yd.sum(0), 1e-6, "conv epilogue: sum of y", atol=1e-6 * (B * T) ** 0.5)
    assert_close_rel(tot[:, 1], (yd * yd).sum(0), 1e-6, "conv epilogue: sum of y^2")
    want = ops.bn_partial_stats(y).sum(0)
    assert_close_rel(tot, want, 1e-6, "against the separate statistics pass", atol=1e-6 * (B * T) ** 0.5)


def test_conv1d_wgrad_many_samples(ops):
    torch.manual_seed(5)
    B, Cin, Cout, T, k = 300, 64, 64, 500, 7
    x = torch.randn(B, Cin, T, device="cuda")
    dy = torch.randn(B, Cout, T, device="cuda")
    dw, _ = ops.conv1d_wgrad(_nwc(dy), _nwc(x), k)
    wd = torch.zeros(Cout, Cin, k, device="cuda", dtype=torch.float64, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv1d(x.double(), wd, None, padding=k // 2), wd, dy.double())
    assert_close_rel(dw, gw, TF32, "conv wgrad B=300")


# ------------------------------------------------------------------ similarity / InfoNCE
@pytest.mark.parametrize("Ml,Ng,D,off", [(128, 128, 128, 0), (256, 256, 128, 0), (200, 600, 64, 200), (4096, 4096, 128, 0),
                                         (5, 5, 32, 0), (1, 1, 16, 0)])
def test_infonce_kernels(ops, Ml, Ng, D, off):
    torch.manual_seed(6)
    e = torch.randn(Ml, D, device="cuda")
    f = torch.randn(Ng, D, device="cuda")
    en, einv = ops.l2norm_fwd(e)
    fn, _ = ops.l2norm_fwd(f)
    # l2norm output is tf32-rounded for the contraction that follows: 2^-11 relative
    assert_close_rel(en, F.normalize(e.double(), dim=1), 5e-4, "l2norm")
    it = 1 / 0.07
    Sref = en.double() @ fn.double().t() * it
    assert_close_rel(ops.similarity(en, fn, it), Sref, TF32, "similarity")
    lse, diag = ops.infonce_lse(en, fn, it, off)
    lref = torch.logsumexp(Sref, dim=1)
    idx = torch.arange(Ml, device="cuda")
    assert float((lse.double() - lref).abs().max()) < 1e-4, "row lse (fp32, |S| <= 14.3)"
    assert float((diag.double() - Sref[idx, idx + off]).abs().max()) < 1e-4
    lse_col = torch.logsumexp(Sref, dim=0).float()
    G = ops.infonce_grad(en, fn, lse, lse_col, it, off, 0.5 / Ml)
    Gref = torch.exp(Sref - lref[:, None]) + torch.exp(Sref - lse_col.double()[None, :])
    Gref[idx, idx + off] -= 2
    assert_close_rel(G, Gref * (0.5 / Ml), TF32, "infonce grad tile", atol=1e-6 / Ml)
    d = torch.randn(Ml, D, device="cuda")
    ed = e.double().requires_grad_(True)
    (gx,) = torch.autograd.grad(F.normalize(ed, dim=1), ed, d.double())
    assert_close_rel(ops.l2norm_bwd(d, en, einv), gx, 5e-4, "l2norm bwd")


@pytest.mark.parametrize("Ml,Ng,off,precise", [(128, 128, 0, True), (256, 256, 0, True), (256, 1024, 512, True),
                                                (4096, 4096, 0, True), (512, 2560, 1024, False), (4096, 4096, 0, False),
                                                (384, 9088, 384, False)])
def test_infonce_bwd_fused(ops, Ml, Ng, off, precise):
    """de = G1 f_n, df = G2 e_n with both gradient blocks formed in tensor memory: against fp64, and the precise mode at
    the accuracy of the unfused 3-pass chain (rows of G sum to ~0: the products cancel against the common component)."""
    torch.manual_seed(11)
    D, it = 128, 1 / 0.07
    e_all = torch.randn(Ng, D, device="cuda") + 0.7
    f_all = e_all * 0.5 + torch.randn(Ng, D, device="cuda") + 0.4
    en_all, e3_all, _ = ops.l2norm_split_fwd(e_all, 0)
    fn_all, f3_all, _ = ops.l2norm_split_fwd(f_all, 1)
    e3, f3 = e3_all[off:off + Ml].contiguous(), f3_all[off:off + Ml].contiguous()
    assert ops.infonce_bwd_fused_supported(Ml, Ng, D, off)
    S = en_all.double() @ fn_all.double().t() * it  # (e rows, f rows) of the global batch
    lse_ef_all, lse_fe_all = torch.logsumexp(S, 1), torch.logsumexp(S, 0)
    coef = 0.5 * it / Ng
    idx = torch.arange(Ml, device="cuda")
    Sl = S[off:off + Ml]                       # my e x all f
    G1 = torch.exp(Sl - lse_ef_all[off:off + Ml, None]) + torch.exp(Sl - lse_fe_all[None, :])
    G1[idx, idx + off] -= 2
    St = S[:, off:off + Ml].t()                # my f x all e
    G2 = torch.exp(St - lse_fe_all[off:off + Ml, None]) + torch.exp(St - lse_ef_all[None, :])
    G2[idx, idx + off] -= 2
    de_ref, df_ref = coef * G1 @ fn_all.double(), coef * G2 @ en_all.double()
    de, df = ops.infonce_bwd_fused(e3, f3, e3_all, f3_all, lse_ef_all[off:off + Ml].float(), lse_fe_all[off:off + Ml].float(),
                                   lse_ef_all.float(), lse_fe_all.float(), it, off, coef, precise)
    l_ef, l_fe, dg = ops.infonce_lse_fused(e3, f3, e3_all, f3_all, it, off)
    assert float((l_ef.double() - lse_ef_all[off:off + Ml]).abs().max()) < 1e-4, "fused row lse, e x f"
    assert float((l_fe.double() - lse_fe_all[off:off + Ml]).abs().max()) < 1e-4, "fused row lse, f x e"
    assert float((dg.double() - S[idx + off, idx + off]).abs().max()) < 1e-4, "fused positives"
    l2 = ops.infonce_lse_fused(e3, f3, e3_all, f3_all, it, off)
    assert torch.equal(l2[0], l_ef) and torch.equal(l2[1], l_fe), "fixed summation order: bit-identical reruns"
    tol = 2e-5 if precise else TF32
    assert_close_rel(de, de_ref, tol, "fused infonce de", atol=1e-9)
    assert_close_rel(df, df_ref, tol, "fused infonce df", atol=1e-9)


@pytest.mark.parametrize("B,H,dh,pdrop", [(37, 4, 32, 0.0), (256, 4, 32, 0.3), (5, 2, 64, 0.0), (3, 1, 16, 0.0), (64, 8, 16, 0.5)])
def test_cross2_attention_of_the_bridge_head(ops, B, H, dh, pdrop):
    """One query over two tokens (bridge_utils.py:74-83) against fp64 autograd; with dropout the kernel's own mask is
    read back from the returned weights (w = mask * p / (1 - drop)) and replayed."""
    torch.manual_seed(21)
    d = H * dh
    q = torch.randn(B, d, device="cuda")
    kv = torch.randn(2 * B, 2 * d, device="cuda")
    dout = torch.randn(B, d, device="cuda")
    out, att = ops.cross2_attn_fwd(q, kv, H, pdrop, 99)
    qd, kvd = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    k = kvd[:, :d].reshape(2, B, H, dh)
    v = kvd[:, d:].reshape(2, B, H, dh)
    p = torch.softmax((qd.reshape(B, H, dh).unsqueeze(0) * k).sum(-1) / dh ** 0.5, dim=0)  # (2, B, H)
    mask = torch.ones_like(p)
    if pdrop > 0:
        mask = (att.permute(2, 0, 1).double() != 0).double() / (1 - pdrop)
        keep = float((mask != 0).double().mean())
        assert abs(keep - (1 - pdrop)) < 4 * (pdrop * (1 - pdrop) / mask.numel()) ** 0.5 + 1e-3
        out2, att2 = ops.cross2_attn_fwd(q, kv, H, pdrop, 99)
        assert torch.equal(att, att2) and torch.equal(out, out2)
        assert not torch.equal(att, ops.cross2_attn_fwd(q, kv, H, pdrop, 100)[1])
    w = p * mask
    ref = (w.unsqueeze(-1) * v).sum(0).reshape(B, d)
    assert_close_rel(out, ref, 1e-5, "cross2 out")
    assert_close_rel(att, w.permute(1, 2, 0), 1e-5, "cross2 weights")
    gq, gkv = torch.autograd.grad(ref, (qd, kvd), dout.double())
    dq, dkv = ops.cross2_attn_bwd(dout, q, kv, H, pdrop, 99)
    assert_close_rel(dq, gq, 1e-5, "cross2 dq", atol=1e-7)
    assert_close_rel(dkv, gkv, 1e-5, "cross2 dkv", atol=1e-7)


# ------------------------------------------------------------------ windowing + band power
@pytest.mark.parametrize("R,C,n,win,hop,nfft,fs", [
    (2, 4, 3000, 1024, 512, 1024, 1000.0), (3, 8, 2000, 500, 250, 512, 250.0), (1, 3, 1001, 100, 37, 128, 128.0),
    (2, 128, 8192, 1024, 512, 1024, 1000.0), (1, 2, 5000, 2000, 1000, 2048, 1000.0), (1, 1, 64, 64, 64, 64, 64.0),
])
def test_window_index_gather_bandpower(ops, R, C, n, win, hop, nfft, fs):
    from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
    from oracle import spectral as osp
    torch.manual_seed(7)
    rec = torch.randn(R, C, n, device="cuda")
    labels = torch.arange(R) % 2
    subj = torch.arange(R) + 100
    st, rid, lab, sub = edu.window_indices(R, n, win, hop, labels, subj)
    o_st, o_rid, o_lab, o_sub = osp.window_indices(R, n, win, hop, labels.numpy(), subj.numpy())
    for got, want in ((st, o_st), (rid, o_rid), (lab, o_lab), (sub, o_sub)):  # bit-exact integer work
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)
    g = edu.gather_windows(rec, win, hop)
    want = osp.gather_windows(rec.cpu().numpy(), win, hop)
    assert np.array_equal(g.cpu().numpy(), want), "gather must be bit-exact"
    gl = edu.gather_windows(rec, win, hop, channels_last=True)
    assert np.array_equal(gl.cpu().numpy(), want.transpose(0, 2, 1))
    p = edu.band_power(rec, fs, win, hop, nfft=nfft)
    ref = osp.band_power(want, fs, nfft=nfft, taper=torch.hann_window(win, periodic=True, dtype=torch.float64).float().double().numpy())
    assert p.shape == ref.shape
    # fp32 FFT vs fp64 oracle: 1e-5 relative per element where the band holds energy
    err = np.abs(p.cpu().numpy().astype(np.float64) - ref) / np.maximum(ref, 1e-30)
    assert err.max() < FP32, f"band power max rel err {err.max():.3e}"


@pytest.mark.parametrize("R,C,n,win,hop,nfft,fs", [
    (2, 128, 8192, 1024, 512, 1024, 1000.0), (1, 64, 4096, 1000, 500, 1024, 1000.0), (2, 130, 4096, 512, 256, 512, 500.0),
    (1, 96, 2048, 256, 128, 256, 250.0), (3, 8, 2000, 500, 248, 500, 500.0), (1, 128, 1024, 1024, 1024, 1024, 1000.0),
    # half-overlapping windows (the overlap kernel): several items per recording (41 windows = 16 + 16 + 9), two windows, one
    (1, 128, 1024 + 40 * 512, 1024, 512, 1024, 1000.0), (2, 32, 1536, 1024, 512, 1024, 1000.0), (2, 64, 1024, 1024, 512, 1024, 1000.0),
    (3, 200, 128 + 16 * 64, 128, 64, 128, 250.0),
])
def test_bandpower_tensor_core_dft(ops, R, C, n, win, hop, nfft, fs):
    """The DFT-as-GEMM band-power kernel (3-pass tf32 with the split samples in tensor memory) against the fp64
    oracle, at the same 1e-5 per-element bound as the FFT kernel, and against the FFT kernel itself."""
    from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
    from oracle import spectral as osp
    torch.manual_seed(8)
    t = torch.arange(n, dtype=torch.float64) / fs
    tones = sum(a * torch.sin(2 * math.pi * f * t + ph) for a, f, ph in ((2.0, 6.0, 0.3), (1.5, 10.0, 1.1), (1.0, 20.0, 2.0)))
    rec = (torch.randn(R, C, n, dtype=torch.float64) + tones).float().cuda()
    p = edu.band_power(rec, fs, win, hop, nfft=nfft, path="dft")
    want = osp.gather_windows(rec.cpu().numpy(), win, hop)
    ref = osp.band_power(want, fs, nfft=nfft, taper=torch.hann_window(win, periodic=True, dtype=torch.float64).float().double().numpy())
    assert p.shape == ref.shape
    err = np.abs(p.cpu().numpy().astype(np.float64) - ref) / np.maximum(ref, 1e-30)
    assert err.max() < FP32, f"DFT band power max rel err {err.max():.3e}"
    if nfft & (nfft - 1) == 0 and nfft >= 64:  # the FFT kernel needs a power-of-two transform length
        pf = edu.band_power(rec, fs, win, hop, nfft=nfft, path="fft")
        assert float(((p - pf).abs() / pf.abs().clamp_min(1e-30)).max()) < 2e-5


def test_bandpower_pure_tone(ops):
    """Known answer: a 10 Hz sinusoid of amplitude A puts A^2/2 in alpha and ~nothing elsewhere."""
    from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
    fs, n = 1000.0, 4096
    t = torch.arange(n, dtype=torch.float64) / fs
    rec = (3.0 * torch.sin(2 * math.pi * 10.0 * t)).float().reshape(1, 1, n).cuda()
    p = edu.band_power(rec, fs, 1000, 1000).cpu()  # nfft 1024, 4 windows
    assert p.shape == (4, 1, 3)
    alpha = p[:, 0, 1]
    assert torch.allclose(alpha, torch.full_like(alpha, 4.5), rtol=2e-2)
    assert float(p[:, 0, 0].max()) < 1e-2 * 4.5 and float(p[:, 0, 2].max()) < 1e-2 * 4.5


def test_window_indices_short_recording(ops):
    from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
    st, rid, lab, sub = edu.window_indices(3, 10, 64, 32)
    assert st.numel() == 0 and rid.numel() == 0 and lab is None and sub is None


# ------------------------------------------------------------------ BN + act (+pool, dropout)
def _bn_ref(y, gamma, beta, act, pool):
    z = F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    a = F.gelu(z) if act == "gelu" else torch.relu(z) if act == "relu" else z
    return F.max_pool1d(a, 2) if pool == 2 else a


@pytest.mark.parametrize("shape,act,pool", [((4, 16, 100), "gelu", 0), ((8, 64, 500), "gelu", 2), ((3, 48, 251), "gelu", 2),
                                            ((64, 128), "relu", 0), ((5, 7, 33), "gelu", 0), ((4096, 64), "relu", 0),
                                            ((2, 128, 250), "none", 0)])
def test_bn_act_fwd_bwd(ops, shape, act, pool):
    torch.manual_seed(8)
    y = torch.randn(*shape, device="cuda") * 2 + 0.5
    C = shape[1]
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda")
    three = y.dim() == 3
    yl = ops.as_nwc(_nwc(y)) if three else y
    cnt = y.numel() // C
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    mean, invstd = ops.bn_finalize_stats(ops.bn_partial_stats(yl), cnt, 1e-5, rm, rv, 0.1)
    out = ops.bn_act_fwd(yl, mean, invstd, gamma, beta, act, pool)
    yd, gd, bd = y.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = _bn_ref(yd, gd, bd, act, pool)
    dout = torch.randn_like(ref).float()
    gy, gg, gb = torch.autograd.grad(ref, (yd, gd, bd), dout.double())
    doutl = ops.as_nwc(_nwc(dout)) if three else dout
    dbeta, dgamma = ops.bn_bwd_finalize(ops.bn_act_bwd_reduce(doutl, yl, mean, invstd, gamma, beta, act, pool))
    dy = ops.bn_act_bwd_apply(doutl, yl, mean, invstd, gamma, beta, dbeta, dgamma, cnt, act, pool)
    t = _nwc if three else (lambda z: z)
    dims = (0, 2) if three else (0,)
    assert_close_rel(out, t(ref), FP32, "bn fwd")
    assert_close_rel(dy, t(gy), FP32, "bn dy")
    assert_close_rel(dgamma, gg, FP32, "bn dgamma")
    assert_close_rel(dbeta, gb, FP32, "bn dbeta")
    assert_close_rel(rm, 0.1 * y.double().mean(dims), FP32, "running mean")
    assert_close_rel(rv, 0.9 + 0.1 * y.double().var(dims, unbiased=True), FP32, "running var")


def test_dropout_mask_statistics_and_consistency(ops):
    torch.manual_seed(9)
    y = ops.as_nwc(torch.randn(8, 500, 64, device="cuda"))
    C = 64
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    mean, invstd = ops.bn_finalize_stats(ops.bn_partial_stats(y), y.numel() // C, 1e-5)
    out = ops.bn_act_fwd(y, mean, invstd, gamma, beta, "none", 2, 0.3, 1234, False)
    out0 = ops.bn_act_fwd(y, mean, invstd, gamma, beta, "none", 2, 0.0, 1234, False)
    keep = out != 0
    assert abs(float(keep.float().mean()) - 0.7) < 5e-3
    assert torch.allclose(out[keep], (out0 / 0.7)[keep], rtol=1e-5)
    # the backward regenerates the same mask from the seed: sum(dz) for dout = 1 counts kept elements / 0.7
    part = ops.bn_act_bwd_reduce(torch.ones_like(out), y, mean, invstd, gamma, beta, "none", 2, 0.3, 1234, False)
    dbeta, _ = ops.bn_bwd_finalize(part)
    assert abs(float(dbeta.sum()) - float(keep.sum()) / 0.7) < 1.0
    a = ops.act_fwd(torch.ones(1 << 16, device="cuda"), "none", 0.4, 77)
    assert abs(float((a != 0).float().mean()) - 0.6) < 1e-2
    da = ops.act_bwd(torch.ones(1 << 16, device="cuda"), torch.ones(1 << 16, device="cuda"), "none", 0.4, 77)
    assert torch.equal(a != 0, da != 0)


def test_seqmean(ops):
    x = ops.as_nwc(torch.randn(6, 250, 96, device="cuda"))
    assert_close_rel(ops.seqmean(x), x.double().mean(1), FP32, "seqmean")
    d = torch.randn(6, 96, device="cuda")
    assert_close_rel(ops.seqmean_bwd(d, 250), (d.double() / 250)[:, None, :].expand(6, 250, 96), FP32, "seqmean bwd")


@pytest.mark.parametrize("M,D,act", [(64, 128, "gelu"), (4096, 128, "gelu"), (100, 64, "relu"), (33, 96, "gelu")])
def test_ln_act(ops, M, D, act):
    torch.manual_seed(10)
    x = torch.randn(M, D, device="cuda") * 1.5 + 0.3
    g = torch.rand(D, device="cuda") + 0.5
    b = torch.randn(D, device="cuda")
    out, mean, rstd = ops.ln_act_fwd(x, g, b, 1e-5, act)
    xd, gd, bd = (t.double().requires_grad_(True) for t in (x, g, b))
    z = F.layer_norm(xd, (D,), gd, bd, 1e-5)
    ref = F.gelu(z) if act == "gelu" else torch.relu(z)
    dout = torch.randn_like(out)
    gx, gg, gb = torch.autograd.grad(ref, (xd, gd, bd), dout.double())
    dx, dg, db = ops.ln_act_bwd(dout, x, g, b, mean, rstd, act)
    assert_close_rel(out, ref, FP32, "ln fwd")
    assert_close_rel(dx, gx, FP32, "ln dx")
    assert_close_rel(dg, gg, FP32, "ln dgamma")
    assert_close_rel(db, gb, FP32, "ln dbeta")


def test_roi_meanstd_zscore_colsum(ops):
    from oracle import models as om
    from oracle import spectral as osp
    torch.manual_seed(11)
    x = torch.randn(64, 100, 200, device="cuda")
    x[0, 3, 5] = float("nan")
    assert_close_rel(ops.roi_meanstd(x), om.roi_meanstd(x.cpu().double()), FP32, "roi mean/std")
    for shape in ((3, 7, 12), (2, 300, 1024), (2, 9, 1028), (5, 33, 203), (4, 1, 8)):  # float4 kernel edge cases, scalar fallback
        xs = torch.randn(*shape, device="cuda") * 2 + 0.5
        assert_close_rel(ops.roi_meanstd(xs), om.roi_meanstd(xs.cpu().double()), FP32, f"roi mean/std {shape}", atol=1e-6)
    x = torch.randn(16, 75, 40, device="cuda") * 3 + 1
    z = ops.zscore(x).cpu()
    for i in range(16):
        assert_close_rel(z[i], osp.normalize_modality(x[i].cpu().double().numpy()), FP32, "normalize_modality")
    x = torch.randn(1000, 96, device="cuda")
    assert_close_rel(ops.colsum(x), x.double().sum(0), FP32, "colsum")


def test_error_codes_surface_as_exceptions(ops):
    from multimodal_eeg_fmri_b200 import _lib
    with pytest.raises(_lib.XmodalError):
        ops.bandpower(torch.randn(1, 1, 100, device="cuda"), 100, 50, 96, 100.0, torch.ones(100, device="cuda"), 100.0,
                      torch.tensor([1, 2], dtype=torch.int32, device="cuda"))  # nfft not a supported power of two
    assert rel_err(torch.ones(2), torch.ones(2)) == 0.0


# ------------------------------------------------------------------ fused attention core
def _attn_ref(qkv, B, L, H, dh, scale, mask=None):
    x = qkv.double().requires_grad_(True)
    q, k, v = (t.reshape(B, L, H, dh).transpose(1, 2) for t in x.chunk(3, dim=-1))
    s = q @ k.transpose(-1, -2) * scale
    p = torch.softmax(s, dim=-1)
    if mask is not None:
        p = p * mask
    return x, s, (p @ v).transpose(1, 2).reshape(B, L, H * dh)


@pytest.mark.parametrize("B,L,H", [(2, 250, 4), (3, 128, 2), (1, 37, 1), (5, 256, 4), (2, 129, 3), (40, 250, 4), (1, 1, 1),
                                   (2, 8, 2), (2, 500, 4), (3, 257, 2), (1, 512, 1), (2, 385, 3)])
def test_fused_attention_fwd_bwd(ops, B, L, H):
    """The on-chip (flash-style) attention core against fp64 torch, dropout off: out, lse, dqkv."""
    torch.manual_seed(14)
    dh = 32
    d = H * dh
    qkv = ops.round_tf32(torch.randn(B, L, 3 * d, device="cuda"))
    dout = ops.round_tf32(torch.randn(B, L, d, device="cuda"))
    scale = dh ** -0.5
    out, lse = ops.attn_fused_fwd(qkv, H, scale, 0.0, 0, round_out=False)
    x, s, ref = _attn_ref(qkv, B, L, H, dh, scale)
    assert_close_rel(out, ref, TF32, "fused attention out")
    assert float((lse.double() - torch.logsumexp(s, -1).reshape(B * H, L)).abs().max()) < 2e-3
    (gx,) = torch.autograd.grad(ref, x, dout.double())
    dqkv = ops.attn_fused_bwd(dout, qkv, out, lse, H, scale, 0.0, 0)
    for name, a, b in zip("qkv", dqkv.chunk(3, dim=-1), gx.chunk(3, dim=-1)):
        assert_close_rel(a, b, 2e-3, f"fused d{name}", atol=1e-6)
    # the in-projection's bias gradient, accumulated inside the kernels: same dqkv, column sums of it (rows past L excluded)
    dqkv2, dbias = ops.attn_fused_bwd(dout, qkv, out, lse, H, scale, 0.0, 0, need_bias=True)
    assert torch.equal(dqkv2, dqkv)
    assert_close_rel(dbias, dqkv.double().sum((0, 1)), 1e-5, "fused in-projection bias gradient", atol=1e-5)


@pytest.mark.parametrize("B,L,H,pdrop", [(2, 250, 4, 0.3), (3, 100, 2, 0.1), (2, 256, 1, 0.5), (2, 500, 2, 0.3)])
def test_fused_attention_dropout(ops, B, L, H, pdrop):
    """Dropout on the attention weights: the exported mask has the right keep rate and no row / column structure,
    forward and both backward kernels regenerate exactly that mask (checked against autograd with it)."""
    torch.manual_seed(15)
    dh, seed = 32, 777
    d = H * dh
    qkv = ops.round_tf32(torch.randn(B, L, 3 * d, device="cuda") * 0.5)
    dout = ops.round_tf32(torch.randn(B, L, d, device="cuda"))
    scale = dh ** -0.5
    keep = ops.attn_fused_mask(B, L, H, pdrop, seed).bool()
    tol = 4 * (pdrop * (1 - pdrop) / keep.numel()) ** 0.5 + 1e-4
    assert abs(float(keep.float().mean()) - (1 - pdrop)) < tol
    # keep rates per query row and per key column scatter like a binomial (no structure along either index)
    sd = (pdrop * (1 - pdrop) / L) ** 0.5
    assert float((keep.float().mean(2) - (1 - pdrop)).abs().max()) < 6 * sd
    assert float((keep.float().mean(1) - (1 - pdrop)).abs().max()) < 6 * sd
    assert not torch.equal(keep, ops.attn_fused_mask(B, L, H, pdrop, seed + 1).bool())
    mask = keep.reshape(B, H, L, L).double() / (1 - pdrop)
    out, lse = ops.attn_fused_fwd(qkv, H, scale, pdrop, seed, round_out=False)
    out2, _ = ops.attn_fused_fwd(qkv, H, scale, pdrop, seed, round_out=False)
    assert torch.equal(out, out2)
    x, s, ref = _attn_ref(qkv, B, L, H, dh, scale, mask)
    assert_close_rel(out, ref, TF32, "fused attention out with dropout")
    assert float((lse.double() - torch.logsumexp(s, -1).reshape(B * H, L)).abs().max()) < 2e-3
    (gx,) = torch.autograd.grad(ref, x, dout.double())
    dqkv = ops.attn_fused_bwd(dout, qkv, out, lse, H, scale, pdrop, seed)
    for name, a, b in zip("qkv", dqkv.chunk(3, dim=-1), gx.chunk(3, dim=-1)):
        assert_close_rel(a, b, 3e-3, f"fused d{name} with dropout")


# ------------------------------------------------------------------ fused feed-forward branch
def _ffn_ref(x, w1, b1, w2, b2, act, mask=None):
    pre = (x.double() @ w1.double().t() + b1.double()).requires_grad_(True)
    a = F.gelu(pre) if act == "gelu" else torch.relu(pre)
    if mask is not None:
        a = a * mask
    return pre, a, a @ w2.double().t() + b2.double()


@pytest.mark.parametrize("M,H,act,pdrop", [(1000, 512, "gelu", 0.0), (70001, 512, "gelu", 0.0), (128, 128, "relu", 0.0),
                                           (4099, 256, "gelu", 0.3), (1, 512, "gelu", 0.0), (20000, 1024, "relu", 0.1),
                                           (148 * 128 * 2 + 5, 512, "gelu", 0.3), (148 * 128 * 3 + 77, 384, "gelu", 0.0),
                                           (148 * 128 * 5, 128, "relu", 0.2)])
def test_ffn_fused_fwd_dgrad(ops, M, H, act, pdrop):
    """linear2(Dropout(act(linear1(x)))) with the hidden activations on chip (enhanced_models_v4.py:79-80,102-105)
    against fp64 torch: y, and from the data-gradient kernel A, dH, dX, db1 -- ragged row counts, 1-8 hidden chunks,
    both activations, dropout replayed through the exported mask.  Outputs carry NaN canaries past row M."""
    torch.manual_seed(21)
    D, seed = 128, 4242
    assert ops.ffn_fused_supported(D, H, act)
    x = ops.round_tf32(torch.randn(M, D, device="cuda"))
    dy = ops.round_tf32(torch.randn(M, D, device="cuda"))
    w1 = ops.round_tf32(torch.randn(H, D, device="cuda") / D ** 0.5)
    w2 = ops.round_tf32(torch.randn(D, H, device="cuda") / H ** 0.5)
    b1, b2 = torch.randn(H, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
    mask = None
    if pdrop > 0:
        keep = ops.ffn_fused_mask(M, H, pdrop, seed).bool()
        assert abs(float(keep.float().mean()) - (1 - pdrop)) < 4 * (pdrop * (1 - pdrop) / keep.numel()) ** 0.5 + 1e-4
        sd = (pdrop * (1 - pdrop) / min(M, H)) ** 0.5
        if M >= 1000:  # no structure along rows or hidden units
            assert float((keep.float().mean(0) - (1 - pdrop)).abs().max()) < 6 * (pdrop * (1 - pdrop) / M) ** 0.5
            assert float((keep.float().mean(1) - (1 - pdrop)).abs().max()) < 6 * (pdrop * (1 - pdrop) / H) ** 0.5
        assert not torch.equal(keep, ops.ffn_fused_mask(M, H, pdrop, seed + 1).bool())
        mask = keep.double() / (1 - pdrop)
    pre, a_ref, y_ref = _ffn_ref(x, w1, b1, w2, b2, act, mask)
    y = ops.ffn_fused_fwd(x, w1, b1, w2, b2, act, pdrop, seed)
    assert torch.equal(y, ops.ffn_fused_fwd(x, w1, b1, w2, b2, act, pdrop, seed))
    assert_close_rel(y, y_ref, TF32, "ffn fused y")
    gpre, = torch.autograd.grad(a_ref, pre, dy.double() @ w2.double(), retain_graph=False)
    a, dh, dx, db1 = ops.ffn_fused_dgrad(x, dy, w1, b1, w2.t().contiguous(), w1.t().contiguous(), act, pdrop, seed)
    assert_close_rel(a, a_ref.detach(), TF32, "ffn fused A")
    assert_close_rel(dh, gpre, TF32, "ffn fused dH")
    assert_close_rel(dx, gpre @ w1.double(), TF32, "ffn fused dX")
    assert_close_rel(db1, gpre.sum(0), TF32, "ffn fused db1", atol=1e-5)


def test_transformer_tail_fused_ffn_matches_unfused(ops, monkeypatch):
    """The whole tail with the fused FFN against the same tail on the linear / act / linear chain (dropout off):
    output and every gradient."""
    from multimodal_eeg_fmri_b200 import functional as XF, modules
    torch.manual_seed(3)
    m = modules.EnhancedERPEncoder(16, 128, 2, 4, 0.0).cuda().train()
    x = torch.randn(6, 16, 120, device="cuda")
    res = {}
    for fused in (True, False):
        monkeypatch.setattr(XF, "_FUSED_FFN", fused)
        m.zero_grad()
        out = m(x)
        out.square().sum().backward()
        res[fused] = (out.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters()})
    assert_close_rel(res[True][0], res[False][0], TF32, "tail output")
    for k, g in res[False][1].items():
        if k.startswith("conv_layers.") and k.endswith(".bias") and not k.startswith("conv_layers.10"):
            continue
        assert_close_rel(res[True][1][k], g, 3e-3, k, atol=1e-6)


# ------------------------------------------------------------------ shape-general attention core
@pytest.mark.parametrize("B,L,H,dh,mask_kind,pdrop", [(3, 50, 4, 8, None, 0.0), (2, 33, 2, 16, "bool2d", 0.0), (2, 40, 4, 32, "float3d", 0.0),
                                                      (1, 700, 2, 64, None, 0.0), (2, 64, 3, 20, "bool3d", 0.0), (2, 96, 4, 8, None, 0.25)])
def test_general_attention_vs_multihead_attention(ops, B, L, H, dh, mask_kind, pdrop):
    """The SIMT attention core (any head dim, attn_mask) against torch fp64 with nn.MultiheadAttention's mask
    semantics: boolean True = NOT allowed to attend, float = additive, (L, L) or (B*H, L, L)."""
    from multimodal_eeg_fmri_b200 import functional as XF
    torch.manual_seed(31)
    d = H * dh
    qkv = torch.randn(B, L, 3 * d, device="cuda")
    dout = torch.randn(B, L, d, device="cuda")
    scale = dh ** -0.5
    mask = None
    if mask_kind == "bool2d":
        mask = torch.triu(torch.ones(L, L, dtype=torch.bool, device="cuda"), 1)  # causal: no fully masked row
    elif mask_kind == "bool3d":
        mask = torch.rand(B * H, L, L, device="cuda") < 0.3
        mask[:, torch.arange(L), torch.arange(L)] = False
    elif mask_kind == "float3d":
        mask = torch.randn(B * H, L, L, device="cuda")
    add = XF.additive_attention_mask(mask, B, H, L)
    x = qkv.double().requires_grad_(True)
    q, k, v = (t.reshape(B, L, H, dh).transpose(1, 2) for t in x.chunk(3, dim=-1))
    s = q @ k.transpose(-1, -2) * scale
    if add is not None:
        s = s + (add.double() if add.dim() == 2 else add.double().reshape(B, H, L, L))
    p = torch.softmax(s, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(B, L, d)
    if pdrop == 0.0:
        out, lse = ops.attn_general_fwd(qkv, H, scale, add, 0.0, 0)
        assert_close_rel(out, ref, FP32 * 3, "general attention out")
        assert float((lse.double() - torch.logsumexp(s, -1).reshape(B * H, L)).abs().max()) < 1e-4
        (gx,) = torch.autograd.grad(ref, x, dout.double())
        dqkv = ops.attn_general_bwd(dout, qkv, lse, H, scale, add, 0.0, 0)
        assert_close_rel(dqkv, gx, 1e-4, "general attention dqkv")
    else:  # dropout: deterministic per seed, keeps the expected mass, backward consistent with the forward (dot test)
        out, lse = ops.attn_general_fwd(qkv, H, scale, add, pdrop, 99)
        out2, _ = ops.attn_general_fwd(qkv, H, scale, add, pdrop, 99)
        assert torch.equal(out, out2) and not torch.equal(out, ops.attn_general_fwd(qkv, H, scale, add, pdrop, 100)[0])
        dqkv = ops.attn_general_bwd(dout, qkv, lse, H, scale, add, pdrop, 99)
        eps = 1e-2
        dirn = torch.randn_like(qkv)
        fp = ops.attn_general_fwd(qkv + eps * dirn, H, scale, add, pdrop, 99)[0]
        fm = ops.attn_general_fwd(qkv - eps * dirn, H, scale, add, pdrop, 99)[0]
        num = float(((fp - fm).double() * dout.double()).sum() / (2 * eps))
        ana = float((dqkv.double() * dirn.double()).sum())
        assert abs(num - ana) <= 2e-2 * max(abs(num), 1.0), (num, ana)


def test_transformer_block_masks_match_torch_multihead_attention():
    """TemporalTransformerBlock.forward(x, mask) against the reference block's composition of torch modules
    (enhanced_models_v4.py:89-107) under the same state_dict: bool / float / per-head masks, head dim 8."""
    from multimodal_eeg_fmri_b200 import modules
    torch.manual_seed(5)
    B, L, d, H = 3, 21, 32, 4
    blk = modules.TemporalTransformerBlock(d, H, 64, 0.0).cuda().train()
    x = torch.randn(B, L, d, device="cuda")

    def reference(x, mask):
        h = blk.norm1(x)
        a, _ = blk.self_attn(h, h, h, attn_mask=mask)
        x = x + a
        return x + blk.linear2(F.gelu(blk.linear1(blk.norm2(x))))

    causal = torch.triu(torch.ones(L, L, dtype=torch.bool, device="cuda"), 1)
    per_head = (torch.rand(B * H, L, L, device="cuda") < 0.4)
    per_head[:, torch.arange(L), torch.arange(L)] = False
    for mask in (None, causal, torch.randn(L, L, device="cuda"), per_head):
        with torch.no_grad():
            ref = reference(x.double().float(), mask)
        assert_close_rel(blk(x, mask), ref, 2e-3, f"block with mask {None if mask is None else (mask.dtype, tuple(mask.shape))}")
    with pytest.raises(ValueError):
        blk(x, torch.zeros(L, L + 1, device="cuda"))


# ------------------------------------------------------------------ ROI connectivity
@pytest.mark.parametrize("B,TR,ROI", [(5, 100, 200), (3, 40, 7), (300, 100, 200), (2, 2, 1), (4, 150, 33)])
def test_roi_corrcoef(ops, B, TR, ROI):
    """Per-sample Pearson correlation matrix of the ROI columns (numpy.corrcoef, NaN -> 0 first) within 1e-5; the
    output buffer carries canaries past its extent; a constant column gives NaN in its row / column like numpy."""
    torch.manual_seed(41)
    x = torch.randn(B, TR, ROI, device="cuda") * 3 + torch.randn(B, 1, ROI, device="cuda") * 10
    if TR > 2:
        x[0, 1, 0] = float("nan")
    assert ops._lib.lib().xm_roi_corrcoef_supported(TR, ROI)
    out = ops.roi_corrcoef(x)
    xc = torch.nan_to_num(x.double())
    xc = xc - xc.mean(1, keepdim=True)
    c = xc.transpose(1, 2) @ xc
    d = torch.sqrt(torch.diagonal(c, dim1=1, dim2=2))
    ref = (c / d[:, :, None] / d[:, None, :]).clamp(-1, 1).reshape(B, -1)
    if TR > 2:
        assert float((out.double() - ref).abs().max()) < 2e-5 and rel_err(out, ref) < FP32
        ref0 = np.corrcoef(np.nan_to_num(x[0].double().cpu().numpy()).T).reshape(-1)
        assert np.abs(out[0].double().cpu().numpy() - ref0).max() < 2e-5
    if ROI >= 7:
        x[1, :, 4] = 1.25
        o = ops.roi_corrcoef(x).reshape(B, ROI, ROI)
        assert torch.isnan(o[1, 4]).all() and torch.isnan(o[1, :, 4]).all() and not torch.isnan(o[1, :4, :4]).any()
        assert not torch.isnan(o[0]).any()


# ------------------------------------------------------------------ producers that emit the tf32 split directly
def test_producers_emit_the_tf32_split(ops):
    """window gather, BatchNorm block and connectivity kernel can write the tf32 split of their output in place of
    a separate xm_split3_f32 pass: bit-identical to splitting the plain output."""
    torch.manual_seed(51)
    rec = torch.randn(3, 20, 700, device="cuda")
    plain = ops.window_gather(rec, 256, 128, channels_last=True)
    got = ops.window_gather(rec, 256, 128, channels_last=True, split3=True)
    G, W, C = plain.shape
    # [hi | lo]: the first two blocks of xm_split3_f32's [hi | lo | hi] (a 3-pass conv wraps its third block onto the first)
    assert got.shape == (G, W, 2 * C) and torch.equal(got.reshape(G * W, 2 * C), ops.split3(plain.reshape(G * W, C), 0, 1)[:, :2 * C])
    y = torch.randn(4, 50, 64, device="cuda")
    mean, invstd = torch.randn(64, device="cuda") * 0.1, torch.rand(64, device="cuda") + 0.5
    gamma, beta = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1
    for pool in (0, 2):
        o = ops.bn_act_fwd(y, mean, invstd, gamma, beta, "gelu", pool)
        o3 = ops.bn_act_fwd(y, mean, invstd, gamma, beta, "gelu", pool, round_out=2)
        assert torch.equal(o3.reshape(-1, 128), ops.split3(o.reshape(-1, 64), 0, 1)[:, :128])
    x = torch.randn(5, 40, 12, device="cuda")
    c = ops.roi_corrcoef(x)
    assert torch.equal(ops.roi_corrcoef(x, prepared=True), ops.linear_precise_prepare(c))


@pytest.mark.parametrize("B,Cin,Cout,T,k", [(3, 64, 64, 500, 7), (2, 64, 128, 250, 5), (2, 32, 48, 100, 3), (2, 20, 32, 64, 5)])
def test_precise_conv_reads_the_two_block_split(ops, B, Cin, Cout, T, k):
    """The 3-pass conv over a producer-written [hi | lo] input (third channel block wrapped onto the first) equals the
    3-pass conv over the explicit [hi | lo | hi] split bit for bit, and is fp32-accurate."""
    torch.manual_seed(8)
    x = torch.randn(B, Cin, T, device="cuda")
    w = torch.randn(Cout, Cin, k, device="cuda") / (Cin * k) ** 0.5
    b = torch.randn(Cout, device="cuda")
    x2 = ops.to_nwc(x, split3=True)  # (B, T, 2 Cin)
    assert x2.shape[2] == 2 * Cin
    y2, xh = ops.conv1d_fwd_precise(x2, w, b)
    y3, _ = ops.conv1d_fwd_precise(ops.to_nwc(x), w, b)
    assert torch.equal(y2, y3)
    assert torch.equal(xh.contiguous(), ops.to_nwc(x, round_out=True))
    ref = _nwc(F.conv1d(x.double(), w.double(), b.double(), padding=k // 2))
    assert_close_rel(y2, ref, 1e-5, "3-pass conv over [hi | lo]")
    y2s, _, part = ops.conv1d_fwd_precise(x2, w, b, stats=True)
    assert torch.equal(y2s, y2)
    assert_close_rel(part.sum(0)[:, 0], y2.double().reshape(-1, Cout).sum(0), 1e-6, "statistics with the wrapped read", atol=1e-6 * (B * T) ** 0.5)


# ------------------------------------------------------------------ out-of-bounds canaries
def test_ragged_outputs_do_not_write_past_their_extent(ops):
    """compute-sanitizer is closed on this pool, so ragged shapes are checked with canaries: outputs are
    carved out of a larger sentinel-filled buffer and the guard bands must stay untouched (the TMA-store
    epilogue has to clip rows/columns that fall outside the tensor; the direct-store path has to guard them)."""
    import ctypes
    from multimodal_eeg_fmri_b200 import _lib
    torch.manual_seed(21)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    SENT = 12345.0
    for (M, N, K) in [(130, 36, 40), (257, 100, 96), (5, 2, 8), (1000, 132, 64)]:
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda")
        guard = 64
        buf = torch.full((guard + M * N + guard,), SENT, device="cuda")
        y = buf[guard:guard + M * N].view(M, N)
        _lib.call("xm_linear_fwd_f32", P(x), P(w), None, P(y), M, N, K, K, K, N, 0, 0, 1, None, st)
        torch.cuda.synchronize()
        assert bool((buf[:guard] == SENT).all()) and bool((buf[guard + M * N:] == SENT).all()), (M, N, K)
        assert_close_rel(y, x.double() @ w.double().t(), TF32, f"linear {M}x{N}x{K}")
    # attention: L not a multiple of 128, outputs embedded in guarded buffers
    B, L, H, dh = 3, 77, 2, 32
    d = H * dh
    qkv = ops.round_tf32(torch.randn(B, L, 3 * d, device="cuda"))
    g = 256
    obuf = torch.full((g + B * L * d + g,), SENT, device="cuda")
    lbuf = torch.full((g + B * H * L + g,), SENT, device="cuda")
    out, lse = obuf[g:-g].view(B, L, d), lbuf[g:-g].view(B * H, L)
    _lib.call("xm_attn_fused_fwd_f32", P(qkv), P(out), P(lse), B, L, H, dh, ctypes.c_float(dh ** -0.5), ctypes.c_float(0.0), 0, 0, st)
    torch.cuda.synchronize()
    for b_ in (obuf, lbuf):
        assert bool((b_[:g] == SENT).all()) and bool((b_[-g:] == SENT).all())
    assert not bool((out == SENT).any()) and not bool((lse == SENT).any())


def test_empty_and_degenerate_inputs(ops):
    """Edge cases of the preprocessing API: recordings shorter than a window, zero recordings, a single
    window, one-sample batches through the BatchNorm-free paths."""
    from multimodal_eeg_fmri_b200 import eeg_data_utils as edu, fmri_utils
    short = torch.randn(2, 4, 100, device="cuda")
    assert edu.band_power(short, 1000.0, 256, 128).shape == (0, 4, 3)
    assert edu.gather_windows(short, 256, 128).shape == (0, 4, 256)
    assert edu.gather_windows(short, 256, 128, channels_last=True).shape == (0, 256, 4)
    st, rid, lab, sub = edu.window_indices(0, 1000, 256, 128)
    assert st.numel() == 0 and rid.numel() == 0
    one = torch.randn(1, 1, 256, device="cuda")
    assert edu.band_power(one, 256.0, 256, 256).shape == (1, 1, 3)  # exactly one window
    assert fmri_utils.aggregate_roi_timeseries(torch.empty(0, 10, 5, device="cuda")).shape == (0, 10)
    x = torch.randn(1, 1, 7, device="cuda")  # a single TR: std = 0
    out = fmri_utils.aggregate_roi_timeseries(x)
    assert torch.equal(out[:, :7], x[:, 0]) and float(out[:, 7:].abs().max()) == 0.0
    e = torch.randn(1, 128, device="cuda", requires_grad=True)
    from multimodal_eeg_fmri_b200.bridge_utils import symmetric_infonce
    loss = symmetric_infonce(e, e.detach().clone())  # batch of one: L = log 1 = 0
    loss.backward()
    assert abs(float(loss)) < 1e-5 and float(e.grad.abs().max()) < 1e-5


@pytest.mark.parametrize("M,C,act,p", [(1000, 512, "gelu", 0.0), (257, 128, "relu", 0.3), (64, 36, "gelu", 0.0)])
def test_act_bwd_with_fused_bias_gradient(ops, M, C, act, p):
    torch.manual_seed(31)
    x = torch.randn(M, C, device="cuda")
    dout = torch.randn(M, C, device="cuda")
    dx, db = ops.act_bwd_colsum(dout, x, act, p, 99)
    ref = ops.act_bwd(dout, x, act, p, 99)  # same mask stream
    assert torch.equal(dx, ref)
    assert_close_rel(db, ref.double().sum(0), FP32, "fused column sums", atol=1e-5)
    if p == 0.0:
        xd = x.double().requires_grad_(True)
        y = F.gelu(xd) if act == "gelu" else torch.relu(xd)
        (g,) = torch.autograd.grad(y, xd, dout.double())
        assert_close_rel(dx, g, FP32, "act backward")
