#!/usr/bin/env python
"""Time the attention core at the bench shape (B=4096, L=250, H=4, dh=32): materialising vs fused kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_eeg_fmri_b200 import ops

B, L, H, dh = int(os.environ.get("AB", 4096)), 250, 4, 32
d = H * dh
torch.manual_seed(0)
qkv = ops.round_tf32(torch.randn(B, L, 3 * d, device="cuda"))
dout = ops.round_tf32(torch.randn(B, L, d, device="cuda"))
scale = dh ** -0.5


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for pd in (0.0, 0.3):
    out, lse = ops.attn_fused_fwd(qkv, H, scale, pd, 7)
    t_ff = timeit(lambda: ops.attn_fused_fwd(qkv, H, scale, pd, 7))
    t_fb = timeit(lambda: ops.attn_fused_bwd(dout, qkv, out, lse, H, scale, pd, 7))
    o2, probs, lse2 = ops.attn_fwd(qkv, H, scale, pd, 7)
    t_of = timeit(lambda: ops.attn_fwd(qkv, H, scale, pd, 7))
    t_ob = timeit(lambda: ops.attn_bwd(dout, qkv, probs, lse2, H, scale, pd, 7))
    del probs
    print(f"p={pd}: fused fwd {t_ff:.3f} ms bwd {t_fb:.3f} ms | materialising fwd {t_of:.3f} ms bwd {t_ob:.3f} ms")
