#!/usr/bin/env python
"""Event-timed fused InfoNCE kernels against the unfused GEMM chain, at the 1-GPU shape (4096 x 4096) and at the per-rank
shape of the 8-GPU weak-scaling run (4096 local rows x 32768 global rows):  python tools/nce_bench.py [--reps 10] [--only-fused]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only-fused", action="store_true")
    args = ap.parse_args()
    torch.manual_seed(0)
    D, it, res = 128, 1 / 0.07, {}
    for Ml, Ng in ((4096, 4096), (4096, 32768)):
        e = torch.randn(Ng, D, device="cuda") + 0.5
        f = e * 0.5 + torch.randn(Ng, D, device="cuda")
        _, e3_all, _ = ops.l2norm_split_fwd(e, 0)
        _, f3_all, _ = ops.l2norm_split_fwd(f, 1)
        e3, f3 = e3_all[:Ml].contiguous(), f3_all[:Ml].contiguous()
        coef = 0.5 * it / Ng
        lse_ef, lse_fe, _ = ops.infonce_lse_fused(e3, f3, e3_all, f3_all, it, 0)
        lse_ef_all = lse_ef.repeat(Ng // Ml)  # timing only: any finite column scales
        lse_fe_all = lse_fe.repeat(Ng // Ml)
        r = {"lse_fused_ms": timed(lambda: ops.infonce_lse_fused(e3, f3, e3_all, f3_all, it, 0), args.reps)}
        for prec in (True, False):
            r["bwd_fused_%s_ms" % ("precise" if prec else "single")] = timed(
                lambda: ops.infonce_bwd_fused(e3, f3, e3_all, f3_all, lse_ef, lse_fe, lse_ef_all, lse_fe_all, it, 0, coef, prec),
                args.reps)
        if not args.only_fused:
            def lse_unfused():
                ops.infonce_lse(e3, f3_all, it, 0)
                ops.infonce_lse(f3, e3_all, it, 0)

            def bwd_unfused(prec):
                G1 = ops.infonce_grad(e3, f3_all, lse_ef, lse_fe_all, it, 0, coef, not prec)
                G2 = ops.infonce_grad(f3, e3_all, lse_fe, lse_ef_all, it, 0, coef, not prec)
                if prec:
                    ops.infonce_dgrad(G1, f3_all, 1)
                    ops.infonce_dgrad(G2, e3_all, 0)
                else:
                    ops.linear_dgrad(G1, f3_all[:, :D])
                    ops.linear_dgrad(G2, e3_all[:, :D])

            r["lse_unfused_ms"] = timed(lse_unfused, args.reps)
            r["bwd_unfused_single_ms"] = timed(lambda: bwd_unfused(False), args.reps)
            if Ng <= 8192:
                r["bwd_unfused_precise_ms"] = timed(lambda: bwd_unfused(True), args.reps)
        flop_s = 2 * 2.0 * Ml * Ng * D * 3  # both directions, 3 score passes
        r["lse_fused_tflops"] = round(flop_s / r["lse_fused_ms"] / 1e9, 1)
        r["bwd_fused_single_tflops"] = round((flop_s + 2 * 2.0 * Ml * Ng * D) / r["bwd_fused_single_ms"] / 1e9, 1)
        res[f"{Ml}x{Ng}"] = {k: round(v, 4) for k, v in r.items()}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
