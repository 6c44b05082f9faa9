"""Phase trace of the fused attention kernels (CTA 0).  Needs a tracing build:
    XM_NVCC_FLAGS=-DXM_FA_TRACE python -m multimodal_eeg_fmri_b200.build --force
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes, torch
from multimodal_eeg_fmri_b200 import ops, _lib
B, L, H, dh = 4096, 250, 4, 32
qkv = ops.round_tf32(torch.randn(B, L, 3 * H * dh, device="cuda"))
ops.attn_fused_fwd(qkv, H, dh ** -0.5, 0.3, 7)
torch.cuda.synchronize()
buf = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda")
_lib.lib().xm_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
ops.attn_fused_fwd(qkv, H, dh ** -0.5, 0.3, 7)
torch.cuda.synchronize()
_lib.lib().xm_debug_set_attn_trace(None)
t = buf.cpu().view(3, 4096)
for role, per in ((0, 4), (1, 5)):
    x = t[role]
    n = int((x != 0).sum())
    x = x[:n]
    print("role", role, "events", n)
    base = int(x[0])
    # steady-state section: items 20..28
    rows = []
    if role == 0:
        ev = x[: (n // 4) * 4].view(-1, 4)
        for i in range(40, 52):
            a = ev[i]
            print("  half", i, "wait_hfull", int(a[1] - a[0]), "issue_mma1", int(a[2] - a[1]), "wait_p_ready", int(a[3] - a[2]), "to_next", int(ev[i + 1][0] - a[3]))
    else:
        # per item: per half 5 stamps (before s_full, after s_full, after ld, after pair barrier, after p_ready) x2 then 2 (before o_full, after o_full)
        ev = x[: (n // 12) * 12].view(-1, 12)
        for i in range(20, 26):
            a = ev[i]
            names = ["h0 wait_s_full", "h0 ld", "h0 max+bar", "h0 exp+st", "h1 pre", "h1 wait_s_full", "h1 ld", "h1 max+bar", "h1 exp+st", "pre_o", "wait_o_full", "to_next"]
            d = [int(a[k + 1] - a[k]) for k in range(11)] + [int(ev[i + 1][0] - a[11])]
            print("  item", i, dict(zip(names, d)), "total", int(ev[i + 1][0] - a[0]))
