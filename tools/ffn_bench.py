#!/usr/bin/env python
"""Event-timed fused FFN kernels at the bench shape (M = 4096 x 250 rows, d_model 128, hidden 512) next to the
unfused chain they replace; also the target of the ncu captures in profiles/ (python tools/ffn_bench.py [--reps N])."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096 * 250)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--drop", type=float, default=0.3)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only-fused", action="store_true")
    a = ap.parse_args()
    M, D, H = a.rows, 128, a.hidden
    torch.manual_seed(0)
    x = ops.round_tf32(torch.randn(M, D, device="cuda"))
    dy = ops.round_tf32(torch.randn(M, D, device="cuda"))
    w1 = ops.round_tf32(torch.randn(H, D, device="cuda") / D ** 0.5)
    w2 = ops.round_tf32(torch.randn(D, H, device="cuda") / H ** 0.5)
    b1, b2 = torch.randn(H, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
    res = {"rows": M, "hidden": H, "drop_p": a.drop}
    res["ffn_fused_fwd_ms"] = timed(lambda: ops.ffn_fused_fwd(x, w1, b1, w2, b2, "gelu", a.drop, 7), a.reps)
    res["ffn_fused_dgrad_ms"] = timed(lambda: ops.ffn_fused_dgrad(x, dy, w1, b1, w2t, w1t, "gelu", a.drop, 7), a.reps)
    res["fwd_tflops"] = 4.0 * M * D * H / res["ffn_fused_fwd_ms"] / 1e9
    res["fwd_algorithmic_gbs"] = 8.0 * M * D / res["ffn_fused_fwd_ms"] / 1e6
    res["dgrad_tflops"] = 6.0 * M * D * H / res["ffn_fused_dgrad_ms"] / 1e9
    res["dgrad_algorithmic_gbs"] = 4.0 * M * (3 * D + 2 * H) / res["ffn_fused_dgrad_ms"] / 1e6
    if not a.only_fused:
        def unfused_fwd():
            f1 = ops.linear_fwd(x, w1, b1)
            g = ops.act_fwd(f1, "gelu", a.drop, 7, round_out=True)
            return ops.linear_fwd(g, w2, b2), f1
        res["unfused_fwd_ms"] = timed(unfused_fwd, a.reps)
        _, f1 = unfused_fwd()

        def unfused_dgrad():
            dg = ops.linear_dgrad(dy, w2)
            df1, db1 = ops.act_bwd_colsum(dg, f1, "gelu", a.drop, 7, round_out=True)
            return ops.linear_dgrad(df1, w1)
        res["unfused_dgrad_ms"] = timed(unfused_dgrad, a.reps)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
