#!/usr/bin/env python
"""Phase trace of the fused FFN forward kernel (CTA 0): where the TMA producer, the MMA issuer and the two transform
groups spend their cycles.  python tools/ffn_trace.py [--rows N] > trace.json"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096 * 250)
    ap.add_argument("--drop", type=float, default=0.3)
    ap.add_argument("--dgrad", action="store_true", help="trace the data-gradient kernel instead of the forward")
    ap.add_argument("--flags", type=int, default=0, help="xm_debug_set_ffn_flags: 1 = no transform arithmetic, 2 = no weight reloads")
    a = ap.parse_args()
    M, D, H = a.rows, 128, 512
    torch.manual_seed(0)
    x = ops.round_tf32(torch.randn(M, D, device="cuda"))
    w1 = ops.round_tf32(torch.randn(H, D, device="cuda") / D ** 0.5)
    w2 = ops.round_tf32(torch.randn(D, H, device="cuda") / H ** 0.5)
    b1, b2 = torch.randn(H, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
    dy = ops.round_tf32(torch.randn(M, D, device="cuda"))
    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
    run = ((lambda: ops.ffn_fused_dgrad(x, dy, w1, b1, w2t, w1t, "gelu", a.drop, 7)) if a.dgrad
           else (lambda: ops.ffn_fused_fwd(x, w1, b1, w2, b2, "gelu", a.drop, 7)))
    for _ in range(2):
        run()
    buf = torch.zeros(3 * 8192, device="cuda", dtype=torch.int64)
    _lib.lib().xm_debug_set_ffn_trace(ctypes.c_void_p(buf.data_ptr()))
    _lib.lib().xm_debug_set_ffn_flags(a.flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    _lib.lib().xm_debug_set_ffn_trace(None)
    _lib.lib().xm_debug_set_ffn_flags(0)
    kernel_ms = e0.elapsed_time(e1)
    t = buf.cpu().numpy()

    def rows(lo, n):
        r = t[lo:lo + n].reshape(-1, 6)
        return r[r[:, 3] != 0]

    prod, mma, g0, g1 = rows(0, 8190), rows(8192, 8190), rows(16384, 4092), rows(16384 + 4096, 4092)
    base = int(min(prod[0, 2], mma[0, 2])) if len(prod) else int(mma[0, 2])
    names = {0: "M1", 1: "M2", 2: "M3", 3: "M4"}
    ops_per_tile = 12 if a.dgrad else 8
    out = {"flags": a.flags, "kernel_ms": kernel_ms, "note": "cycles relative to the first event of CTA 0; wait_w = cycles blocked on the weight ring, wait_o = on other roles"}
    out["mma"] = [dict(op=names[int(r[0])], n=int(r[1]), start=int(r[2]) - base, end=int(r[3]) - base, wait_w=int(r[4]), wait_o=int(r[5]))
                  for r in mma[:48]]
    out["producer"] = [dict(op=names[int(r[0])], n=int(r[1]), start=int(r[2]) - base, end=int(r[3]) - base, wait_w_empty=int(r[4]),
                            wait_x_empty=int(r[5])) for r in prod[:48]]
    tf = lambda r: dict(n=int(r[1]), wait_start=int(r[2]) - base, h_full=int(r[3]) - base, arrived=int(r[4]) - base,
                        stored=(int(r[5]) - base if r[5] else None))
    out["transform_g0"] = [tf(r) for r in g0[:24]]
    out["transform_g1"] = [tf(r) for r in g1[:24]]
    n_t = len(mma) // ops_per_tile
    if n_t > 4:  # steady state summary over the CTA's tiles
        m = mma[2 * ops_per_tile:]
        per = len(m) / float(ops_per_tile)
        out["steady_state"] = {"ops": len(m), "cycles_per_tile": float((m[-1, 3] - m[0, 2]) / per),
                               "mma_wait_w_per_tile": float(m[:, 4].sum() / per),
                               "mma_wait_other_per_tile": float(m[:, 5].sum() / per),
                               "transform_busy_per_chunk_g0": float((g0[4:, 4] - g0[4:, 3]).mean()),
                               "transform_wait_per_chunk_g0": float((g0[4:, 3] - g0[4:, 2]).mean()),
                               "transform_busy_per_chunk_g1": float((g1[4:, 4] - g1[4:, 3]).mean()) if len(g1) > 4 else None,
                               "transform_wait_per_chunk_g1": float((g1[4:, 3] - g1[4:, 2]).mean()) if len(g1) > 4 else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
