"""Phase trace of the fused attention kernels (CTA 0).  Needs a tracing build:
    XM_NVCC_FLAGS=-DXM_FA_TRACE python -m multimodal_eeg_fmri_b200.build --force
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes, torch
from multimodal_eeg_fmri_b200 import ops, _lib
B, L, H, dh = 4096, 250, 4, 32
qkv = ops.round_tf32(torch.randn(B, L, 3 * H * dh, device="cuda"))
dout = ops.round_tf32(torch.randn(B, L, H * dh, device="cuda"))
out, lse = ops.attn_fused_fwd(qkv, H, dh ** -0.5, 0.3, 7)
ops.attn_fused_bwd(dout, qkv, out, lse, H, dh ** -0.5, 0.3, 7)
torch.cuda.synchronize()
# the two backward kernels share the buffer: run them separately by tracing, then reading after each call is not
# possible (one C call) -> the KV kernel (second) overwrites the dq kernel's stamps; XM_FA_TRACE_KERNEL picks one
buf = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda")
_lib.lib().xm_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
ops.attn_fused_bwd(dout, qkv, out, lse, H, dh ** -0.5, 0.3, 7)
torch.cuda.synchronize()
_lib.lib().xm_debug_set_attn_trace(None)
t = buf.cpu().view(3, 4096)
x = t[0]; n = int((x != 0).sum()); x = x[:n]
# MMA warp: per item 1 + 4 chunks x 4 stamps = 17
ev = x[: (n // 17) * 17].view(-1, 17)
for i in range(20, 24):
    a = ev[i]
    d = [int(a[k + 1] - a[k]) for k in range(16)] + [int(ev[i + 1][0] - a[16])]
    print("mma item", i, "wait_item", d[0], [dict(wait_hfull=d[1 + 4 * c], issue_s=d[2 + 4 * c], wait_p=d[3 + 4 * c], out_mma_next=d[4 + 4 * c]) for c in range(4)], "total", int(ev[i + 1][0] - a[0]))
x = t[1]; n = int((x != 0).sum()); x = x[:n]
# epilogue warp 2: per item 1 + 4 x 3 + 1 = 14
ev = x[: (n // 14) * 14].view(-1, 14)
for i in range(20, 24):
    a = ev[i]
    d = [int(a[k + 1] - a[k]) for k in range(13)] + [int(ev[i + 1][0] - a[13])]
    print("epi item", i, "pre", d[0], [dict(wait_s=d[1 + 3 * c], work=d[2 + 3 * c], gap=d[3 + 3 * c]) for c in range(4)], "tail", d[13], "total", int(ev[i + 1][0] - a[0]))
