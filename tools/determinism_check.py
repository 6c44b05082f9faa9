#!/usr/bin/env python
"""Bitwise repeatability of the step's pieces: every op / block is run twice on the same inputs and the results compared
with torch.equal (max abs difference printed where they differ).  python tools/determinism_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops, synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel  # noqa: E402


def cmp(name, a, b):
    a = a if isinstance(a, (tuple, list)) else (a,)
    b = b if isinstance(b, (tuple, list)) else (b,)
    bad = []
    for i, (x, y) in enumerate(zip(a, b)):
        if x is None or not torch.is_tensor(x):
            continue
        if not torch.equal(x, y):
            d = (x.double() - y.double()).abs().max().item()
            bad.append((i, d, d / (y.double().abs().max().item() + 1e-30)))
    print(f"{name:40s} {'identical' if not bad else 'DIFFERS ' + str(bad)}", flush=True)


def twice(name, fn):
    a = fn()
    torch.cuda.synchronize()
    b = fn()
    torch.cuda.synchronize()
    cmp(name, a, b)


torch.manual_seed(0)
B, L, D, Hh = 512, 250, 128, 512
M = B * L
x = ops.round_tf32(torch.randn(M, D, device="cuda"))
dy = ops.round_tf32(torch.randn(M, D, device="cuda"))
w1 = ops.round_tf32(torch.randn(Hh, D, device="cuda") / D ** 0.5)
w2 = ops.round_tf32(torch.randn(D, Hh, device="cuda") / Hh ** 0.5)
b1, b2 = torch.randn(Hh, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
twice("ffn_fused_fwd", lambda: ops.ffn_fused_fwd(x, w1, b1, w2, b2, "gelu", 0.3, 7))
twice("ffn_fused_dgrad", lambda: ops.ffn_fused_dgrad(x, dy, w1, b1, w2.t().contiguous(), w1.t().contiguous(), "gelu", 0.3, 7))
qkv = ops.round_tf32(torch.randn(B, L, 3 * D, device="cuda"))
dout = ops.round_tf32(torch.randn(B, L, D, device="cuda"))
out, lse = ops.attn_fused_fwd(qkv, 4, 32 ** -0.5, 0.3, 5)
twice("attn_fused_fwd", lambda: ops.attn_fused_fwd(qkv, 4, 32 ** -0.5, 0.3, 5))
twice("attn_fused_bwd (+bias)", lambda: ops.attn_fused_bwd(dout, qkv, out, lse, 4, 32 ** -0.5, 0.3, 5, round_out=True, need_bias=True))
twice("linear_fwd", lambda: ops.linear_fwd(x, w1, b1))
twice("linear_wgrad", lambda: ops.linear_wgrad(dy, x))
twice("linear_dgrad", lambda: ops.linear_dgrad(dy, w2.t().contiguous()))
xc = torch.randn(B, 500, 64, device="cuda")
wc = torch.randn(64, 64, 7, device="cuda") / 21
wk, wt = ops.conv1d_pack_weight(wc)
bc = torch.randn(64, device="cuda")
twice("conv1d_fwd (+stats)", lambda: ops.conv1d_fwd(xc, wk, bc, 64, stats=True))
twice("conv1d_fwd_precise", lambda: ops.conv1d_fwd_precise(xc, wc, bc))
dyc = torch.randn(B, 500, 64, device="cuda")
twice("conv1d_dgrad", lambda: ops.conv1d_dgrad(dyc, wt, 64))
twice("conv1d_wgrad", lambda: ops.conv1d_wgrad(dyc, xc, 7))
e = torch.randn(2048, D, device="cuda") + 0.5
f = e * 0.5 + torch.randn(2048, D, device="cuda")
_, e3, _ = ops.l2norm_split_fwd(e, 0)
_, f3, _ = ops.l2norm_split_fwd(f, 1)
le, lf, _ = ops.infonce_lse_fused(e3, f3, e3, f3, 1 / 0.07, 0)
twice("infonce_lse_fused", lambda: ops.infonce_lse_fused(e3, f3, e3, f3, 1 / 0.07, 0))
twice("infonce_bwd_fused precise", lambda: ops.infonce_bwd_fused(e3, f3, e3, f3, le, lf, le, lf, 1 / 0.07, 0, 1e-3, True))
twice("infonce_bwd_fused single", lambda: ops.infonce_bwd_fused(e3, f3, e3, f3, le, lf, le, lf, 1 / 0.07, 0, 1e-3, False))
roi = torch.randn(256, 100, 200, device="cuda")
twice("roi_corrcoef", lambda: ops.roi_corrcoef(roi))

# the whole model: gradients of two identical forward / backward passes
torch.manual_seed(42)
m = PairedBridgeModel(64, 200, None, 128, 64, 128, 0.0, 0.0, "v4").cuda().train()
eeg, roi_s, conn = (t.cuda() for t in synthetic.paired_batch(1024, 64, 500, 200, 100, seed=42))


def grads(overlap):
    m.overlap_branches = overlap
    for p in m.parameters():
        p.grad = None
    loss = m(eeg, roi_s, conn)
    loss.backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, loss.detach().clone()


for overlap in (False, True):
    g1, l1 = grads(overlap)
    g2, l2 = grads(overlap)
    diff = {k: ((g1[k].double() - g2[k].double()).norm() / (g2[k].double().norm() + 1e-30)).item() for k in g1 if not torch.equal(g1[k], g2[k])}
    worst = sorted(diff.items(), key=lambda kv: -kv[1])[:6]
    print(f"model grads, overlap_branches={overlap}: loss identical {bool(torch.equal(l1, l2))}; {len(diff)} of {len(g1)} tensors differ; worst {worst}", flush=True)
