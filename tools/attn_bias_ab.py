#!/usr/bin/env python
"""A/B: fused attention backward with and without the in-kernel in-projection bias gradient (batch 4096, L 250, 4 heads)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops  # noqa: E402

B, L, H, dh = 4096, 250, 4, 32
torch.manual_seed(0)
qkv = ops.round_tf32(torch.randn(B, L, 3 * H * dh, device="cuda"))
dout = ops.round_tf32(torch.randn(B, L, H * dh, device="cuda"))
out, lse = ops.attn_fused_fwd(qkv, H, dh ** -0.5, 0.3, 5)
res = {}
for nb in (False, True, False, True):
    for _ in range(2):
        ops.attn_fused_bwd(dout, qkv, out, lse, H, dh ** -0.5, 0.3, 5, round_out=True, need_bias=nb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attn_fused_bwd(dout, qkv, out, lse, H, dh ** -0.5, 0.3, 5, round_out=True, need_bias=nb)
    e1.record()
    torch.cuda.synchronize()
    res.setdefault("need_bias" if nb else "plain", []).append(round(e0.elapsed_time(e1) / 5, 4))
print(json.dumps(res))
