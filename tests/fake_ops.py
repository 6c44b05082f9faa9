"""TEST-ONLY fake backend: torch-CPU stand-ins for the raw device ops of multimodal_eeg_fmri_b200.ops, so
the host-side logic above the C ABI (autograd functions, drop-in modules, SyncBN / all-gather /
gradient-bucket plumbing of the data-parallel step) can be exercised without a GPU, including
world_size-2 `gloo` runs.  Installed by monkeypatching attributes of the real `ops` module inside a
test; never imported by the product package (whose ops raise on non-CUDA tensors).

Each fake follows the contract documented in include/xmodal_b200.h for the entry point it replaces
(fp64 internally, fp32 out; dropout supported only with p = 0)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

_ACT = {None: "none", 0: "none", 1: "relu", 2: "gelu", 3: "tanh", 4: "sigmoid"}


def _act(z, act):
    act = _ACT.get(act, act)
    return {"none": lambda t: t, "relu": torch.relu, "gelu": F.gelu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](z)


def _nodrop(p):
    assert float(p) == 0.0, "the fake backend has no dropout mask generator (parity runs use dropout = 0)"


def as_nwc(t):
    return t.contiguous()


def empty_pitched(shape, device):
    return torch.empty(*shape, device=device, dtype=torch.float32)


def _stack3(t):
    """The fake's channel-stacked 'split' [hi | lo] of an exact value: hi = the value, lo = 0."""
    return torch.cat([t, torch.zeros_like(t)], dim=-1).contiguous()


def window_gather(rec, win, hop, channels_last=False, round_out=False, split3=False):
    R, C, n = rec.shape
    w = rec.unfold(2, win, hop).permute(0, 2, 1, 3).reshape(-1, C, win)  # (R*n_win, C, win)
    w = (w.transpose(1, 2) if channels_last else w).contiguous()
    return _stack3(w) if split3 else w


def to_nwc(x, round_out=False, split3=False):
    y = x.transpose(1, 2).contiguous()
    return _stack3(y) if split3 else y


# ---------------------------------------------------------------- linear
def linear_fwd(x, w, bias=None, act=None, round_out=False, splits=0):
    return _act(F.linear(x.double(), w.double(), None if bias is None else bias.double()), act).float()


def linear_dgrad(dy, w, round_out=False):
    return (dy.double() @ w.double()).float()


def linear_wgrad(dy, x, need_bias=True, splits=0):
    return (dy.double().t() @ x.double()).float(), (dy.double().sum(0).float() if need_bias else None)


# ---------------------------------------------------------------- conv1d on channels-last activations
def conv1d_pack_weight(w):
    return w.detach().clone(), w.detach().clone()  # both "packed" forms are just the weight here


def conv1d_fwd(x, wk, bias, Cout, round_out=False, out=None, stats=False, wrap_cin=0):
    y = F.conv1d(x.double().transpose(1, 2), wk.double(), None if bias is None else bias.double(), padding=wk.shape[-1] // 2)
    y = y.transpose(1, 2).float().contiguous()
    if out is not None:
        out.copy_(y)
        y = out
    return (y, bn_partial_stats(y)) if stats else y


def conv1d_fwd_precise(x, w, bias, stats=False):
    if x.shape[2] == 2 * w.shape[1]:  # channel-stacked split [hi | lo] from the producer
        x = x[:, :, : w.shape[1]]
    if stats:
        y, part = conv1d_fwd(x, w, bias, w.shape[0], stats=True)
        return y, x, part
    return conv1d_fwd(x, w, bias, w.shape[0]), x


def conv1d_dgrad(dy, wt, Cin, round_out=False, out=None):
    dx = F.conv_transpose1d(dy.double().transpose(1, 2), wt.double(), padding=wt.shape[-1] // 2)
    dx = dx.transpose(1, 2).float().contiguous()
    if out is not None:
        out.copy_(dx)
        return out
    return dx


@torch.enable_grad()
def conv1d_wgrad(dy, x, taps, need_bias=True):
    Cout, Cin = dy.shape[2], x.shape[2]
    w = torch.zeros(Cout, Cin, taps, dtype=torch.float64, requires_grad=True)
    y = F.conv1d(x.double().transpose(1, 2), w, None, padding=taps // 2)
    (gw,) = torch.autograd.grad(y, w, dy.double().transpose(1, 2))
    return gw.float(), (dy.double().sum((0, 1)).float() if need_bias else None)


# ---------------------------------------------------------------- batch norm + act (+pool)
def _rows(y):
    return y.reshape(-1, y.shape[-1])


def bn_partial_stats(y):
    r = _rows(y).double()
    return torch.stack([r.sum(0), (r * r).sum(0)], dim=1).unsqueeze(0)  # (1, C, 2)


def bn_finalize_stats(part, count, eps, running_mean=None, running_var=None, momentum=0.1):
    s = part.sum(0)
    mean = s[:, 0] / count
    var = (s[:, 1] / count - mean * mean).clamp_min(0)
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean.float())
        running_var.mul_(1 - momentum).add_(momentum * (var * count / max(count - 1, 1)).float())
    return mean.float(), torch.rsqrt(var + eps).float()


def _bn_formula(y, mean, invstd, gamma, beta, act, pool):
    z = (y - mean) * invstd * gamma + beta
    a = _act(z, act)
    if pool == 2:
        B, T, C = a.shape
        a = a[:, : T // 2 * 2].reshape(B, T // 2, 2, C).amax(2)
    return z, a


def bn_act_fwd(y, mean, invstd, gamma, beta, act, pool=0, drop_p=0.0, seed=0, drop_before_pool=False, round_out=False):
    _nodrop(drop_p)
    _, a = _bn_formula(y.double(), mean.double(), invstd.double(), gamma.double(), beta.double(), act, pool)
    a = a.float().contiguous()
    return _stack3(a) if int(round_out) == 2 else a


@torch.enable_grad()
def _dz(dout, y, mean, invstd, gamma, beta, act, pool):
    z0 = ((y.double() - mean.double()) * invstd.double() * gamma.double() + beta.double()).requires_grad_(True)
    a = _act(z0, act)
    if pool == 2:
        B, T, C = a.shape
        a = a[:, : T // 2 * 2].reshape(B, T // 2, 2, C).amax(2)
    (dz,) = torch.autograd.grad(a, z0, dout[..., : y.shape[-1]].double())  # a 2C-wide dout: gradient in the first block
    xhat = (y.double() - mean.double()) * invstd.double()
    return dz, xhat


def bn_act_bwd_reduce(dout, y, mean, invstd, gamma, beta, act, pool=0, drop_p=0.0, seed=0, drop_before_pool=False):
    _nodrop(drop_p)
    dz, xhat = _dz(dout, y, mean, invstd, gamma, beta, act, pool)
    return torch.stack([_rows(dz).sum(0), _rows(dz * xhat).sum(0)], dim=1).unsqueeze(0)


def bn_bwd_finalize(part):
    s = part.sum(0)
    return s[:, 0].float(), s[:, 1].float()


def bn_act_bwd_apply(dout, y, mean, invstd, gamma, beta, dbeta, dgamma, count, act, pool=0, drop_p=0.0, seed=0,
                     drop_before_pool=False, round_out=False):
    dz, xhat = _dz(dout, y, mean, invstd, gamma, beta, act, pool)
    dy = gamma.double() * invstd.double() * (dz - dbeta.double() / count - xhat * dgamma.double() / count)
    return dy.float().contiguous()


def seqmean(x):
    return x.double().mean(1).float()


def seqmean_bwd(dout, T):
    return (dout.double() / T).unsqueeze(1).expand(dout.shape[0], T, dout.shape[1]).float().contiguous()


# ---------------------------------------------------------------- layer norm + act, act, reductions
def ln_act_fwd(x, gamma, beta, eps, act, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    xd = x.double()
    mean = xd.mean(1)
    rstd = torch.rsqrt(xd.var(1, unbiased=False) + eps)
    z = (xd - mean[:, None]) * rstd[:, None] * gamma.double() + beta.double()
    return _act(z, act).float(), mean.float(), rstd.float()


@torch.enable_grad()
def ln_act_bwd(dout, x, gamma, beta, mean, rstd, act, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    xd, gd, bd = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    eps = 0.0  # rebuild from the saved statistics: identical to layer_norm's own backward
    m = xd.mean(1, keepdim=True)
    v = xd.var(1, unbiased=False, keepdim=True)
    saved_eps = (1.0 / rstd.double() ** 2 - v.detach().squeeze(1)).clamp_min(0).mean()
    z = (xd - m) * torch.rsqrt(v + saved_eps + eps) * gd + bd
    gx, gg, gb = torch.autograd.grad(_act(z, act), (xd, gd, bd), dout.double())
    return gx.float(), gg.float(), gb.float()


def act_fwd(x, act, drop_p=0.0, seed=0, round_out=False):
    _nodrop(drop_p)
    return _act(x.double(), act).float()


@torch.enable_grad()
def act_bwd(dout, x, act, drop_p=0.0, seed=0, round_out=False):
    _nodrop(drop_p)
    xd = x.double().requires_grad_(True)
    (g,) = torch.autograd.grad(_act(xd, act), xd, dout.double())
    return g.float()


def round_tf32(x, inplace=False):
    return x


def linear_fwd_precise(x, w, bias=None, act=None):
    return linear_fwd(x, w, bias, act)


def linear_dgrad_precise(dy, w):
    return linear_dgrad(dy, w)


def linear_wgrad_precise(dy, x, need_bias=True):
    return linear_wgrad(dy, x, need_bias)


def linear_precise_prepare(x):
    return x


def linear_fwd_prepared(x3, w, bias=None, act=None):
    return linear_fwd(x3, w, bias, act)


def linear_wgrad_prepared(dy, x3, need_bias=True, K=None):
    return linear_wgrad(dy, x3, need_bias)


def act_bwd_colsum(dout, x, act, drop_p=0.0, seed=0, round_out=False):
    dx = act_bwd(dout, x, act, drop_p, seed, round_out)
    return dx, dx.double().sum(0).float()


def colsum(x):
    return x.double().sum(0).float()


# ---------------------------------------------------------------- l2norm / similarity / InfoNCE
def l2norm_fwd(x, eps=1e-12):
    n = x.double().norm(dim=1).clamp_min(eps)
    return (x.double() / n[:, None]).float(), (1.0 / n).float()


def l2norm_split_fwd(x, which, eps=1e-12):
    xn, inv = l2norm_fwd(x, eps)
    z = torch.zeros_like(xn)
    return xn, torch.cat([xn, z, z], 1), inv


def infonce_dgrad(G, f3, which_f):
    D = f3.shape[1] // 3
    return (G.double() @ f3[:, :D].double()).float()  # the fake split keeps the full value in block 0


def l2norm_bwd(dxn, xn, inv):
    d, x = dxn.double(), xn.double()
    return ((d - x * (x * d).sum(1, keepdim=True)) * inv.double()[:, None]).float()


def similarity(a, b, inv_tau):
    return (a.double() @ b.double().t() * inv_tau).float()


def infonce_lse(a, b, inv_tau, diag_off=0):
    S = a.double() @ b.double().t() * inv_tau
    i = torch.arange(a.shape[0])
    return torch.logsumexp(S, 1).float(), S[i, i + diag_off].float()


def infonce_grad(a, b, lse_row, lse_col, inv_tau, diag_off, coef, round_out=False):
    S = a.double() @ b.double().t() * inv_tau
    G = torch.exp(S - lse_row.double()[:, None]) + torch.exp(S - lse_col.double()[None, :])
    i = torch.arange(a.shape[0])
    G[i, i + diag_off] -= 2.0
    return (G * coef).float()


def infonce_bwd_fused_supported(Ml, Ng, D, diag_off):
    return D == 128 and Ml % 128 == 0 and Ng % 128 == 0 and diag_off % 128 == 0 and diag_off + Ml <= Ng


def infonce_lse_fused(e3, f3, e3_all, f3_all, inv_tau, diag_off=0):
    lse_ef, diag = infonce_lse(e3, f3_all, inv_tau, diag_off)
    lse_fe, _ = infonce_lse(f3, e3_all, inv_tau, diag_off)
    return lse_ef, lse_fe, diag


def infonce_bwd_fused(e3, f3, e3_all, f3_all, lse_ef, lse_fe, lse_ef_all, lse_fe_all, inv_tau, diag_off, coef, precise=True):
    D = e3.shape[1] // 3
    G1 = infonce_grad(e3, f3_all, lse_ef, lse_fe_all, inv_tau, diag_off, coef)
    G2 = infonce_grad(f3, e3_all, lse_fe, lse_ef_all, inv_tau, diag_off, coef)
    return (G1.double() @ f3_all[:, :D].double()).float(), (G2.double() @ e3_all[:, :D].double()).float()


# ---------------------------------------------------------------- attention core
def attn_supported(L, dh):
    return dh == 32 and 0 < L <= 512


def _attn_parts(qkv, nhead):
    B, L, E = qkv.shape
    d = E // 3
    q, k, v = (t.reshape(B, L, nhead, d // nhead).transpose(1, 2).double() for t in qkv.chunk(3, dim=-1))
    return q, k, v


def _attn_fwd(qkv, nhead, scale, drop_p=0.0, seed=0, round_out=True):
    _nodrop(drop_p)
    B, L, E = qkv.shape
    q, k, v = _attn_parts(qkv, nhead)
    s = q @ k.transpose(-1, -2) * scale
    p = torch.softmax(s, dim=-1)
    out = (p @ v).transpose(1, 2).reshape(B, L, E // 3)
    return out.float(), p.reshape(B * nhead, L, L).float(), torch.logsumexp(s, -1).reshape(B * nhead, L).float()


def _attn_bwd(dout, qkv, probs, lse, nhead, scale, drop_p=0.0, seed=0, round_out=False):
    _nodrop(drop_p)
    B, L, E = qkv.shape
    d = E // 3
    q, k, v = _attn_parts(qkv, nhead)
    p = probs.reshape(B, nhead, L, L).double()
    do = dout.reshape(B, L, nhead, d // nhead).transpose(1, 2).double()
    dv = p.transpose(-1, -2) @ do
    dp = do @ v.transpose(-1, -2)
    ds = p * (dp - (p * dp).sum(-1, keepdim=True)) * scale
    dq, dk = ds @ k, ds.transpose(-1, -2) @ q
    back = lambda t: t.transpose(1, 2).reshape(B, L, d)
    return torch.cat([back(dq), back(dk), back(dv)], dim=-1).float()


def attn_fused_fwd(qkv, nhead, scale, drop_p=0.0, seed=0, round_out=True):
    out, _, lse = _attn_fwd(qkv, nhead, scale, drop_p, seed, round_out)
    return out, lse


def attn_fused_bwd(dout, qkv, out, lse, nhead, scale, drop_p=0.0, seed=0, round_out=False, need_bias=False):
    _nodrop(drop_p)
    B, L, E = qkv.shape
    q, k, _ = _attn_parts(qkv, nhead)
    probs = torch.exp(q @ k.transpose(-1, -2) * scale - lse.reshape(B, nhead, L, 1).double())
    dqkv = _attn_bwd(dout, qkv, probs.reshape(B * nhead, L, L), lse, nhead, scale, drop_p, seed, round_out)
    return (dqkv, dqkv.double().sum((0, 1)).float()) if need_bias else dqkv


def attn_general_supported(L, dh):
    return 0 < dh <= 256 and L > 0


def _attn_general_scores(qkv, nhead, scale, mask):
    B, L, E = qkv.shape
    q, k, v = _attn_parts(qkv, nhead)
    s = q @ k.transpose(-1, -2) * scale
    if mask is not None:
        s = s + (mask.double() if mask.dim() == 2 else mask.double().reshape(B, nhead, L, L))
    return q, k, v, s


def attn_general_fwd(qkv, nhead, scale, mask=None, drop_p=0.0, seed=0, round_out=False):
    _nodrop(drop_p)
    B, L, E = qkv.shape
    q, k, v, s = _attn_general_scores(qkv, nhead, scale, mask)
    out = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, E // 3)
    return out.float(), torch.logsumexp(s, -1).reshape(B * nhead, L).float()


def attn_general_bwd(dout, qkv, lse, nhead, scale, mask=None, drop_p=0.0, seed=0, round_out=False):
    _nodrop(drop_p)
    B, L, E = qkv.shape
    q, k, v, s = _attn_general_scores(qkv, nhead, scale, mask)
    probs = torch.exp(s - lse.reshape(B, nhead, L, 1).double())
    return _attn_bwd(dout, qkv, probs.reshape(B * nhead, L, L), lse, nhead, scale, drop_p, seed, round_out)


# ---------------------------------------------------------------- fused feed-forward branch
def ffn_fused_supported(D, hidden, act):
    return D == 128 and hidden % 128 == 0 and 0 < hidden <= 1024 and _ACT.get(act, act) in ("gelu", "relu")


def ffn_fused_fwd(x, w1, b1, w2, b2, act, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    h = _act(F.linear(x.double(), w1.double(), b1.double()), act)
    return F.linear(h, w2.double(), b2.double()).float()


def ffn_fused_dgrad(x, dy, w1, b1, w2t, w1t, act, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    pre = F.linear(x.double(), w1.double(), b1.double()).requires_grad_(True)
    with torch.enable_grad():
        a = _act(pre, act)
    (dh,) = torch.autograd.grad(a, pre, dy.double() @ w2t.double().t())
    return a.detach().float(), dh.float(), (dh @ w1t.double().t()).float(), dh.sum(0).float()


# ---------------------------------------------------------------- residual stream (transformer block)
def resid_ln_supported(D):
    return D % 128 == 0 and 128 <= D <= 512


def resid_ln_fwd(x, a, gamma, beta, eps, drop_p=0.0, seed=0, pe=None, L=0):
    _nodrop(drop_p)
    s = x.double()
    if pe is not None:
        s = s + pe.double().repeat(x.shape[0] // L, 1)
    if a is not None:
        s = s + a.double()
    mean = s.mean(1)
    rstd = torch.rsqrt(s.var(1, unbiased=False) + eps)
    h = (s - mean[:, None]) * rstd[:, None] * gamma.double() + beta.double()
    return (x if (a is None and pe is None) else s.float()), h.float(), mean.float(), rstd.float()


def resid_ln_bwd(dh, dres, s, gamma, mean, rstd, drop_p=0.0, seed=0, need_da=True):
    _nodrop(drop_p)
    xh = (s.double() - mean.double()[:, None]) * rstd.double()[:, None]
    dz = dh.double() * gamma.double()
    ds = rstd.double()[:, None] * (dz - dz.mean(1, keepdim=True) - xh * (dz * xh).mean(1, keepdim=True))
    if dres is not None:
        ds = ds + dres.double()
    return (ds.float(), (ds.float() if need_da else None), (dh.double() * xh).sum(0).float(), dh.double().sum(0).float(),
            (ds.sum(0).float() if need_da else None))


def resid_seqmean_fwd(x, a, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    s = x.double() + (0 if a is None else a.double())
    return s.mean(1).float()


def resid_seqmean_bwd(dout, T, drop_p=0.0, seed=0, need_dx=True, need_da=True):
    _nodrop(drop_p)
    g = (dout.double() / T).unsqueeze(1).expand(dout.shape[0], T, dout.shape[1]).float().contiguous()
    return (g if need_dx else None), (g.clone() if need_da else None)


# ---------------------------------------------------------------- preprocessing
def roi_meanstd(x):
    xd = torch.nan_to_num(x.double(), nan=0.0)
    return torch.cat([xd.mean(1), xd.std(1, unbiased=False)], dim=1).float()


def roi_corrcoef(x, prepared=False):  # the fake's "prepared" split is the identity (linear_precise_prepare above)
    xc = torch.nan_to_num(x.double())
    xc = xc - xc.mean(1, keepdim=True)
    c = xc.transpose(1, 2) @ xc
    d = torch.sqrt(torch.diagonal(c, dim1=1, dim2=2))
    return (c / d[:, :, None] / d[:, None, :]).clamp(-1, 1).reshape(x.shape[0], -1).float()


def clip_adamw_(p, g, m, v, step, lr, weight_decay, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8):
    norm = g.double().norm().float()
    if max_norm > 0:
        g.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
    p.mul_(1 - lr * weight_decay)
    m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    c1, c2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    p.addcdiv_(m, v.sqrt() / c2 ** 0.5 + eps, value=-lr / c1)
    return norm.reshape(1)


def zscore(x, eps=1e-8):
    xd = x.double().reshape(x.shape[0], -1)
    out = (xd - xd.mean(1, keepdim=True)) / (xd.std(1, unbiased=False, keepdim=True) + eps)
    return out.reshape(x.shape).float()


def _cross2_parts(q, kv, nhead):
    B, d = q.shape
    dh = d // nhead
    qd = q.double().reshape(B, nhead, dh)
    k = kv[:, :d].double().reshape(2, B, nhead, dh)
    v = kv[:, d:].double().reshape(2, B, nhead, dh)
    p = torch.softmax((qd.unsqueeze(0) * k).sum(-1) / dh ** 0.5, dim=0)  # (2, B, H)
    return qd, k, v, p


def cross2_attn_fwd(q, kv, nhead, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    qd, k, v, p = _cross2_parts(q, kv, nhead)
    out = (p.unsqueeze(-1) * v).sum(0).reshape(q.shape)
    return out.float(), p.permute(1, 2, 0).contiguous().float()


@torch.enable_grad()
def cross2_attn_bwd(dout, q, kv, nhead, drop_p=0.0, seed=0):
    _nodrop(drop_p)
    qd, kvd = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    B, d = q.shape
    dh = d // nhead
    k = kvd[:, :d].reshape(2, B, nhead, dh)
    v = kvd[:, d:].reshape(2, B, nhead, dh)
    p = torch.softmax((qd.reshape(B, nhead, dh).unsqueeze(0) * k).sum(-1) / dh ** 0.5, dim=0)
    out = (p.unsqueeze(-1) * v).sum(0).reshape(B, d)
    dq, dkv = torch.autograd.grad(out, (qd, kvd), dout.double())
    return dq.float(), dkv.float()


def gather_flat_(dst, tensors, offsets):
    for t, o in zip(tensors, offsets):
        dst[o:o + t.numel()].copy_(t.reshape(-1))
    return dst


FAKES = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in ("F", "torch", "annotations")]


def install(monkeypatch=None):
    """Replace the raw device ops with the fakes above (attributes of the real module)."""
    from multimodal_eeg_fmri_b200 import ops

    for name in FAKES:
        if name == "install":
            continue
        if monkeypatch is not None:
            monkeypatch.setattr(ops, name, globals()[name])
        else:
            setattr(ops, name, globals()[name])
