"""Kernel bring-up on a real B200: each group runs in its own subprocess (a trapped kernel leaves
a sticky CUDA error) and prints relative errors against fp64 torch references.

    python tests/gpu_bringup.py            # all groups
    python tests/gpu_bringup.py conv_fwd   # one group, in-process

Not collected by pytest (the parity tests proper are tests/test_gpu_*.py); this is the fast
diagnostic used while developing kernels.
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    import torch
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def g_linear_fwd():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(0)
    for (M, N, K, act, splits) in [(128, 64, 64, None, 1), (300, 96, 200, "gelu", 1), (4096, 128, 400, "relu", 1),
                                   (256, 128, 4096, None, 4), (70, 2, 32, None, 1), (512, 256, 128, None, 1),
                                   (1000, 320, 96, None, 1)]:
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda")
        y = ops.linear_fwd(x, w, b, act=act, splits=splits)
        ref = x.double() @ w.double().t() + b.double()
        if act == "gelu":
            ref = torch.nn.functional.gelu(ref)
        if act == "relu":
            ref = torch.relu(ref)
        print(f"linear_fwd M{M} N{N} K{K} act={act} splits={splits}: rel={rel(y, ref):.3e}", flush=True)


def g_linear_dgrad():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(1)
    for (M, N, K) in [(128, 32, 32), (300, 96, 200), (4096, 128, 400), (256, 64, 128), (512, 2, 64)]:
        dy = torch.randn(M, N, device="cuda")
        w = torch.randn(N, K, device="cuda")
        dx = ops.linear_dgrad(dy, w)
        ref = dy.double() @ w.double()
        print(f"linear_dgrad M{M} N{N} K{K}: rel={rel(dx, ref):.3e}", flush=True)


def g_linear_wgrad():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(2)
    for (M, N, K, splits) in [(128, 128, 32, 1), (300, 96, 200, 1), (4096, 128, 400, 4), (256, 64, 128, 1),
                              (2048, 128, 4000, 2), (512, 2, 64, 1)]:
        dy = torch.randn(M, N, device="cuda")
        x = torch.randn(M, K, device="cuda")
        dw, db = ops.linear_wgrad(dy, x, splits=splits)
        ref = dy.double().t() @ x.double()
        print(f"linear_wgrad M{M} N{N} K{K} splits={splits}: rel={rel(dw, ref):.3e} db={rel(db, dy.double().sum(0)):.3e}",
              flush=True)


def _conv_cases():
    sel = os.environ.get("XM_CASE")
    cases = _conv_cases_all()
    return cases if sel is None else [cases[int(sel)]]


def _conv_cases_all():
    return [(1, 32, 32, 128, 1), (2, 32, 32, 128, 1), (2, 32, 32, 128, 3), (3, 64, 64, 500, 7), (2, 64, 128, 500, 5), (2, 128, 128, 250, 3), (2, 48, 96, 250, 5),
            (2, 18, 48, 500, 7), (1, 192, 128, 100, 1)]


def _nwc(x):
    return x.transpose(1, 2).contiguous()


def g_conv_fwd():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(3)
    for (B, Cin, Cout, T, k) in _conv_cases():
        x = torch.randn(B, Cin, T, device="cuda")
        w = torch.randn(Cout, Cin, k, device="cuda") / (Cin * k) ** 0.5
        b = torch.randn(Cout, device="cuda")
        wk, wt = ops.conv1d_pack_weight(w)
        xl = ops.to_nwc(x)
        ok_t = bool((xl == _nwc(x)).all())
        y = ops.conv1d_fwd(xl, wk, b, Cout)
        ref = torch.nn.functional.conv1d(x.double(), w.double(), b.double(), padding=k // 2)
        print(f"conv_fwd B{B} Cin{Cin} Cout{Cout} T{T} k{k}: rel={rel(y, _nwc(ref)):.3e} to_nwc_exact={ok_t}", flush=True)


def g_conv_dgrad():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(4)
    for (B, Cin, Cout, T, k) in _conv_cases():
        dy = torch.randn(B, Cout, T, device="cuda")
        w = torch.randn(Cout, Cin, k, device="cuda") / (Cin * k) ** 0.5
        wk, wt = ops.conv1d_pack_weight(w)
        dx = ops.conv1d_dgrad(_nwc(dy), wt, Cin)
        ref = torch.nn.functional.conv_transpose1d(dy.double(), w.double(), padding=k // 2)
        print(f"conv_dgrad B{B} Cin{Cin} Cout{Cout} T{T} k{k}: rel={rel(dx, _nwc(ref)):.3e}", flush=True)


def g_conv_wgrad():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(5)
    for (B, Cin, Cout, T, k) in _conv_cases() + ([(300, 64, 64, 500, 7)] if os.environ.get("XM_CASE") is None else []):
        x = torch.randn(B, Cin, T, device="cuda")
        dy = torch.randn(B, Cout, T, device="cuda")
        dw, db = ops.conv1d_wgrad(_nwc(dy), _nwc(x), k)
        xd = x.double().requires_grad_(False)
        wd = torch.zeros(Cout, Cin, k, device="cuda", dtype=torch.float64, requires_grad=True)
        out = torch.nn.functional.conv1d(xd, wd, None, padding=k // 2)
        (gw,) = torch.autograd.grad(out, wd, dy.double())
        print(f"conv_wgrad B{B} Cin{Cin} Cout{Cout} T{T} k{k}: rel={rel(dw, gw):.3e} db={rel(db, dy.double().sum((0, 2))):.3e}",
              flush=True)


def g_infonce():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(6)
    for (Ml, Ng, D, off) in [(128, 128, 128, 0), (256, 256, 128, 0), (200, 600, 64, 200), (4096, 4096, 128, 0)]:
        e = torch.randn(Ml, D, device="cuda")
        f = torch.randn(Ng, D, device="cuda")
        en, einv = ops.l2norm_fwd(e)
        fn, finv = ops.l2norm_fwd(f)
        ref_en = torch.nn.functional.normalize(e.double(), dim=1)
        print(f"l2norm rel={rel(en, ref_en):.3e}")
        it = 1 / 0.07
        S = ops.similarity(en, fn, it)
        Sref = en.double() @ fn.double().t() * it
        lse, diag = ops.infonce_lse(en, fn, it, off)
        lref = torch.logsumexp(Sref, dim=1)
        dref = Sref[torch.arange(Ml), torch.arange(Ml) + off]
        lse_col = torch.logsumexp(Sref, dim=0).float()
        G = ops.infonce_grad(en, fn, lse, lse_col, it, off, 0.5 / Ml)
        Gref = (torch.exp(Sref - lref[:, None]) + torch.exp(Sref - lse_col.double()[None, :]))
        Gref[torch.arange(Ml), torch.arange(Ml) + off] -= 2
        Gref *= 0.5 / Ml
        print(f"infonce Ml{Ml} Ng{Ng} D{D}: S rel={rel(S, Sref):.3e} lse maxabs={float((lse.double()-lref).abs().max()):.3e} "
              f"diag maxabs={float((diag.double()-dref).abs().max()):.3e} G rel={rel(G, Gref):.3e}", flush=True)
        d = torch.randn(Ml, D, device="cuda")
        dx = ops.l2norm_bwd(d, en, einv)
        ed = e.double().requires_grad_(True)
        (gx,) = torch.autograd.grad(torch.nn.functional.normalize(ed, dim=1), ed, d.double())
        print(f"l2norm_bwd rel={rel(dx, gx):.3e}", flush=True)


def g_bandpower():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(7)
    for (R, C, n, win, hop, nfft, fs) in [(2, 4, 3000, 1024, 512, 1024, 1000.0), (3, 8, 2000, 500, 250, 512, 250.0),
                                          (1, 3, 1001, 100, 37, 128, 128.0), (2, 128, 8192, 1024, 512, 1024, 1000.0),
                                          (1, 2, 5000, 2000, 1000, 2048, 1000.0)]:
        rec = torch.randn(R, C, n, device="cuda")
        taper = torch.hann_window(win, periodic=True, device="cuda", dtype=torch.float64)
        bands = [(4, 8), (8, 13), (13, 30)]
        import math
        bins = []
        for lo, hi in bands:
            bins += [math.ceil(lo * nfft / fs), math.ceil(hi * nfft / fs)]
        bins_t = torch.tensor(bins, device="cuda", dtype=torch.int32)
        tp32 = taper.float().contiguous()
        sumsq = float((tp32.double() ** 2).sum())
        p = ops.bandpower(rec, win, hop, nfft, fs, tp32, sumsq, bins_t)
        wins = rec.double().unfold(2, win, hop)  # (R, C, n_win, win)
        X = torch.fft.rfft(wins * tp32.double(), n=nfft)
        P = X.abs() ** 2
        P[..., 1:nfft // 2] *= 2
        P = P / (nfft * sumsq)
        ref = torch.stack([P[..., bins[2 * i]:bins[2 * i + 1]].sum(-1) for i in range(3)], -1)  # (R,C,nw,3)
        ref = ref.permute(0, 2, 1, 3).reshape(-1, C, 3)
        print(f"bandpower R{R} C{C} n{n} win{win} hop{hop} nfft{nfft}: rel={rel(p, ref):.3e} "
              f"maxrel={float(((p.double()-ref).abs()/ref).max()):.3e}", flush=True)
        st, rid, lab, sub = ops.window_index(R, n, win, hop, torch.arange(R, device='cuda') % 2,
                                             torch.arange(R, device='cuda') + 100)
        nw = (n - win) // hop + 1
        ok = bool((st == (torch.arange(R * nw, device='cuda') % nw) * hop).all() and
                  (rid == torch.arange(R * nw, device='cuda') // nw).all() and (lab == rid % 2).all() and
                  (sub == rid + 100).all())
        g = ops.window_gather(rec, win, hop)
        okg = bool((g == rec.unfold(2, win, hop).permute(0, 2, 1, 3).reshape(-1, C, win)).all())
        print(f"  window_index exact={ok} gather exact={okg}", flush=True)


def _bn_ref(y, gamma, beta, act, pool, eps=1e-5):
    import torch
    F = torch.nn.functional
    z = F.batch_norm(y, None, None, gamma, beta, True, 0.1, eps)
    a = F.gelu(z) if act == "gelu" else torch.relu(z) if act == "relu" else z
    if pool == 2:
        a = F.max_pool1d(a, 2)
    return a


def g_bn():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(8)
    for (shape, act, pool) in [((4, 16, 100), "gelu", 0), ((8, 64, 500), "gelu", 2), ((3, 48, 251), "gelu", 2),
                               ((64, 128), "relu", 0), ((5, 7, 33), "gelu", 0), ((4096, 64), "relu", 0)]:
        y = torch.randn(*shape, device="cuda") * 2 + 0.5   # reference layout (B, C, T) or (B, C)
        C = shape[1]
        gamma = torch.rand(C, device="cuda") + 0.5
        beta = torch.randn(C, device="cuda")
        three = y.dim() == 3
        yl = ops.as_nwc(_nwc(y)) if three else y
        part = ops.bn_partial_stats(yl)
        cnt = y.numel() // C
        rm = torch.zeros(C, device="cuda")
        rv = torch.ones(C, device="cuda")
        mean, invstd = ops.bn_finalize_stats(part, cnt, 1e-5, rm, rv, 0.1)
        out = ops.bn_act_fwd(yl, mean, invstd, gamma, beta, act, pool)
        yd = y.double().requires_grad_(True)
        gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
        ref = _bn_ref(yd, gd, bd, act, pool)
        dout = torch.randn_like(ref).float()
        gy, gg, gb = torch.autograd.grad(ref, (yd, gd, bd), dout.double())
        doutl = ops.as_nwc(_nwc(dout)) if three else dout
        part2 = ops.bn_act_bwd_reduce(doutl, yl, mean, invstd, gamma, beta, act, pool)
        dbeta, dgamma = ops.bn_bwd_finalize(part2)
        dy = ops.bn_act_bwd_apply(doutl, yl, mean, invstd, gamma, beta, dbeta, dgamma, cnt, act, pool)
        dims = (0, 2) if three else (0,)
        t = (lambda z: _nwc(z)) if three else (lambda z: z)
        print(f"bn {shape} act={act} pool={pool}: fwd rel={rel(out, t(ref)):.3e} dy rel={rel(dy, t(gy)):.3e} "
              f"dgamma rel={rel(dgamma, gg):.3e} dbeta rel={rel(dbeta, gb):.3e} "
              f"rmean rel={rel(rm, 0.1 * y.double().mean(dims)):.3e} rvar rel={rel(rv, 0.9 + 0.1 * y.double().var(dims, unbiased=True)):.3e}",
              flush=True)
    # dropout statistics + fwd/bwd mask consistency
    y = ops.as_nwc(torch.randn(8, 500, 64, device="cuda"))
    C = 64
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    mean, invstd = ops.bn_finalize_stats(ops.bn_partial_stats(y), y.numel() // C, 1e-5)
    for dbp in (False, True):
        out = ops.bn_act_fwd(y, mean, invstd, gamma, beta, "none", 2, 0.3, 1234, dbp)
        out0 = ops.bn_act_fwd(y, mean, invstd, gamma, beta, "none", 2, 0.0, 1234, dbp)
        if not dbp:
            keep = (out != 0).float().mean().item()
            ok = torch.allclose(out[out != 0], (out0 / 0.7)[out != 0], rtol=1e-5)
            print(f"dropout(after pool) keep={keep:.4f} (want 0.7) scaled_ok={ok}", flush=True)
        dout = torch.ones_like(out)
        part2 = ops.bn_act_bwd_reduce(dout, y, mean, invstd, gamma, beta, "none", 2, 0.3, 1234, dbp)
        dbeta, _ = ops.bn_bwd_finalize(part2)
        expect = float((out != 0).sum()) / 0.7 if not dbp else None
        print(f"dropout dbp={dbp} sum(dz)={float(dbeta.sum()):.1f} expect={expect}", flush=True)
    x = ops.as_nwc(torch.randn(6, 250, 96, device="cuda"))
    print(f"seqmean rel={rel(ops.seqmean(x), x.double().mean(1)):.3e}")
    d = torch.randn(6, 96, device="cuda")
    print(f"seqmean_bwd rel={rel(ops.seqmean_bwd(d, 250), (d.double() / 250)[:, None, :].expand(6, 250, 96)):.3e}", flush=True)


def g_ln():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(9)
    F = torch.nn.functional
    for (M, D, act) in [(64, 128, "gelu"), (4096, 128, "gelu"), (100, 64, "relu"), (33, 96, "gelu")]:
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
        print(f"ln M{M} D{D} {act}: fwd={rel(out, ref):.3e} dx={rel(dx, gx):.3e} dg={rel(dg, gg):.3e} db={rel(db, gb):.3e}",
              flush=True)


def g_misc():
    import torch
    from multimodal_eeg_fmri_b200 import ops
    torch.manual_seed(10)
    x = torch.randn(64, 100, 200, device="cuda")
    x[0, 3, 5] = float("nan")
    out = ops.roi_meanstd(x)
    xd = torch.nan_to_num(x.double(), nan=0.0)
    ref = torch.cat([xd.mean(1), xd.std(1, unbiased=False)], 1)
    print(f"roi_meanstd rel={rel(out, ref):.3e}")
    x = torch.randn(16, 75, 40, device="cuda") * 3 + 1
    z = ops.zscore(x)
    xd = x.double()
    ref = (xd - xd.mean((1, 2), keepdim=True)) / (xd.std((1, 2), unbiased=False, keepdim=True) + 1e-8)
    print(f"zscore rel={rel(z, ref):.3e}")
    x = torch.randn(1000, 96, device="cuda")
    print(f"colsum rel={rel(ops.colsum(x), x.double().sum(0)):.3e}")


def g_tma_probe():
    """Which TMA coordinates are legal?  Each probe runs in its own subprocess via XM_CASE."""
    import ctypes
    import torch
    from multimodal_eeg_fmri_b200 import _lib
    cases = [(0, 0, 0), (4, 0, 0), (-4, 0, 0), (0, -1, 0), (0, 3, 0), (1, 0, 0), (-1, 0, 0), (2, 0, 0), (1, 0, 1), (0, 0, 1), (-4, 0, 1)]
    c0, c1, sw = cases[int(os.environ.get("XM_CASE", "0"))]
    rows, cols = 64, 96
    src = torch.arange(rows * cols, device="cuda", dtype=torch.float32).reshape(rows, cols) + 1
    out = torch.zeros(1024, device="cuda")
    _lib.call("xm_debug_tma_probe", ctypes.c_void_p(src.data_ptr()), rows, cols, cols, c0, c1, sw,
              ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    img = out.cpu().reshape(32, 32)
    # un-swizzle: SWIZZLE_128B: 16-B chunk ^= row % 8 ; ATOM_32B: 32-B chunk ^= row % 4
    exp = torch.zeros(32, 32)
    for r in range(32):
        for c in range(32):
            rr, cc = c1 + r, c0 + c
            exp[r, c] = float(src[rr, cc]) if (0 <= rr < rows and 0 <= cc < cols) else 0.0
    un = torch.zeros(32, 32)
    for r in range(32):
        for c in range(32):
            if sw == 0:
                pc = (((c // 4) ^ (r % 8)) * 4) + (c % 4)
            else:
                pc = (((c // 8) ^ (r % 4)) * 8) + (c % 8)
            un[r, c] = img[r, pc]
    print(f"tma_probe c0={c0} c1={c1} atom32={sw}: match={bool((un == exp).all())}", flush=True)


GROUPS = {k[2:]: v for k, v in list(globals().items()) if k.startswith("g_")}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        GROUPS[sys.argv[1]]()
        import torch
        torch.cuda.synchronize()
        print(f"[{sys.argv[1]}] done", flush=True)
        sys.exit(0)
    rc = 0
    runs = []
    only = os.environ.get("XM_GROUPS")
    for name in GROUPS:
        if only and name not in only.split(","):
            continue
        if name == "tma_probe":
            runs += [(name, str(i)) for i in range(11)]
        elif name.startswith("conv_") and os.environ.get("XM_SPLIT_CONV"):
            runs += [(name, str(i)) for i in range(len(_conv_cases_all()))]
        else:
            runs.append((name, None))
    for name, case in runs:
        t0 = time.time()
        env = dict(os.environ)
        if case is not None:
            env["XM_CASE"] = case
            env["XM_SYNC_DEBUG"] = "1"
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=240, env=env)
            print(r.stdout, end="")
            if r.returncode != 0:
                rc = 1
                print(f"[{name} case={case}] FAILED rc={r.returncode}\n{r.stderr[-1500:]}")
        except subprocess.TimeoutExpired as e:
            rc = 1
            print(f"[{name}] TIMEOUT\n{(e.stdout or b'')[-2000:]}")
        print(f"[{name}] {time.time() - t0:.1f}s", flush=True)
    sys.exit(rc)
