#!/usr/bin/env python
"""Per-entry-point micro-benchmark at the bench.py shapes (B = 4096 paired step): CUDA-event times with
an L2 flush between launches, achieved TFLOP/s and GB/s from the algorithmic work annotated in ops.py.

    python tools/kbench.py [pattern ...] [--batch 4096] [--reps 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import ops  # noqa: E402


def cases(B):
    T, L = 500, 250
    M = B * L
    r = lambda *s: torch.randn(*s, device="cuda")
    c = {}
    # conv stack of the v4 ERP encoder (channels-last)
    for name, (cin, cout, t, k) in {"conv1": (64, 64, T, 7), "conv2": (64, 128, T, 5), "conv3": (128, 128, L, 3)}.items():
        def mk(cin=cin, cout=cout, t=t, k=k):
            x, dy, w = r(B, t, cin), r(B, t, cout), r(cout, cin, k)
            wk, wt = ops.conv1d_pack_weight(w)
            b = r(cout)
            return {"fwd": lambda: ops.conv1d_fwd(x, wk, b, cout), "dgrad": lambda: ops.conv1d_dgrad(dy, wt, cin),
                    "wgrad": lambda: ops.conv1d_wgrad(dy, x, k)}
        c[name] = mk
    # transformer-tail projections over (B*L, d) rows
    for name, (n, k) in {"qkv": (384, 128), "proj": (128, 128), "ffn1": (512, 128), "ffn2": (128, 512)}.items():
        def mk(n=n, k=k):
            x, dy, w, b = r(M, k), r(M, n), r(n, k), r(n)
            return {"fwd": lambda: ops.linear_fwd(x, w, b), "dgrad": lambda: ops.linear_dgrad(dy, w),
                    "wgrad": lambda: ops.linear_wgrad(dy, x)}
        c["lin_" + name] = mk
    def fmri():
        x, w, b, dy = r(B, 40000), r(128, 40000), r(128), r(B, 128)
        return {"fwd": lambda: ops.linear_fwd(x, w, b), "wgrad": lambda: ops.linear_wgrad(dy, x)}
    c["lin_conn"] = fmri
    def bn():
        y = r(B, T, 64)
        g, b = r(64), r(64)
        part = ops.bn_partial_stats(y)
        mean, invstd = ops.bn_finalize_stats(part, B * T, 1e-5)
        out = ops.bn_act_fwd(y, mean, invstd, g, b, "gelu")
        dout = torch.randn_like(out)
        p2 = ops.bn_act_bwd_reduce(dout, y, mean, invstd, g, b, "gelu")
        db, dg = ops.bn_bwd_finalize(p2)
        return {"stats": lambda: ops.bn_partial_stats(y), "fwd": lambda: ops.bn_act_fwd(y, mean, invstd, g, b, "gelu"),
                "bwd_reduce": lambda: ops.bn_act_bwd_reduce(dout, y, mean, invstd, g, b, "gelu"),
                "bwd_apply": lambda: ops.bn_act_bwd_apply(dout, y, mean, invstd, g, b, db, dg, B * T, "gelu")}
    c["bn64"] = bn
    def nce():
        e, f = ops.l2norm_fwd(r(B, 128))[0], ops.l2norm_fwd(r(B, 128))[0]
        lse, _ = ops.infonce_lse(e, f, 1 / 0.07)
        return {"lse": lambda: ops.infonce_lse(e, f, 1 / 0.07), "grad": lambda: ops.infonce_grad(e, f, lse, lse, 1 / 0.07, 0, 1e-4)}
    c["infonce"] = nce
    def attn():
        qkv = ops.round_tf32(r(B, L, 384))
        dout = ops.round_tf32(r(B, L, 128))
        out, probs, lse = ops.attn_fwd(qkv, 4, 32 ** -0.5, 0.3, 77)
        return {"fwd": lambda: ops.attn_fwd(qkv, 4, 32 ** -0.5, 0.3, 77),
                "bwd": lambda: ops.attn_bwd(dout, qkv, probs, lse, 4, 32 ** -0.5, 0.3, 77)}
    c["attn"] = attn

    def attn_fused():
        qkv = ops.round_tf32(torch.randn(B, 250, 384, device="cuda"))
        dout = ops.round_tf32(torch.randn(B, 250, 128, device="cuda"))
        out, lse = ops.attn_fused_fwd(qkv, 4, 32 ** -0.5, 0.3, 77)
        return {"fwd": lambda: ops.attn_fused_fwd(qkv, 4, 32 ** -0.5, 0.3, 77),
                "bwd": lambda: ops.attn_fused_bwd(dout, qkv, out, lse, 4, 32 ** -0.5, 0.3, 77)}
    c["attn_fused"] = attn_fused
    def pre():
        rec = r(max(B // 64, 1), 128, 512 * 65)
        from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
        x = r(B, 64, T)
        roi = r(B, 100, 200)
        return {"bandpower": lambda: edu.band_power(rec, 1000.0, 1024, 512), "to_nwc": lambda: ops.to_nwc(x, True),
                "roi_meanstd": lambda: ops.roi_meanstd(roi)}
    c["pre"] = pre
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("patterns", nargs="*")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"{'case':28s} {'ms':>9s} {'TFLOP/s':>9s} {'GB/s':>9s}")
    for group, mk in cases(a.batch).items():
        if a.patterns and not any(p in group for p in a.patterns):
            continue
        fns = mk()
        for name, fn in fns.items():
            fn()
            torch.cuda.synchronize()
            tot, work = 0.0, (0.0, 0.0)
            for _ in range(a.reps):
                flush.zero_()
                ops.start_timeline()
                fn()
                tl = ops.stop_timeline()
                tot += sum(v[1] for v in tl.values())
                work = (sum(v[2] for v in tl.values()), sum(v[3] for v in tl.values()))
            ms = tot / a.reps
            print(f"{group + '.' + name:28s} {ms:9.4f} {work[0] / ms / 1e9:9.2f} {work[1] / ms / 1e6:9.1f}", flush=True)
        del fns
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
