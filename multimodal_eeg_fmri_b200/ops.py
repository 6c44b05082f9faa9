"""Raw (non-autograd) device ops: thin tensor-level wrappers over the C ABI.

Every function takes CUDA fp32 tensors, allocates its outputs/workspaces with torch (device
memory plumbing only), enqueues the hand-written kernels on torch's current stream and returns.
No function here has a CPU path: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
_ACT = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU, "tanh": ACT_TANH, "sigmoid": ACT_SIGMOID}

_launches = 0  # C-ABI calls made (bench.py reports it as gpu_launches lower bound)


def launch_count() -> int:
    return _launches


def act_code(act) -> int:
    return act if isinstance(act, int) else _ACT[act]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.XmodalError("xmodal-b200 ops need CUDA tensors (no CPU fallback on this path)")
        if t.dtype != torch.float32:
            raise _lib.XmodalError(f"expected float32, got {t.dtype}")


_SYNC_DEBUG = bool(int(__import__("os").environ.get("XM_SYNC_DEBUG", "0")))


_timeline = None  # when a list: (entry point, start event, end event) per C-ABI call (bench.py's per-kernel timing)


def start_timeline() -> None:
    global _timeline
    _timeline = []


def stop_timeline(raw: bool = False):
    """-> {entry point: (calls, total ms, algorithmic flops, algorithmic bytes)}; times are CUDA events on the
    launching stream around each C-ABI call.  raw=True: the calls in launch order, [(entry point, ms, flops, bytes)]."""
    global _timeline
    tl, _timeline = _timeline or [], None
    torch.cuda.synchronize()
    if raw:
        return [(name, a.elapsed_time(b), fl, by) for name, a, b, (fl, by) in tl]
    out = {}
    for name, a, b, (fl, by) in tl:
        n, ms, f0, b0 = out.get(name, (0, 0.0, 0.0, 0.0))
        out[name] = (n + 1, ms + a.elapsed_time(b), f0 + fl, b0 + by)
    return out


_work = (0.0, 0.0)


def _w(flops: float, nbytes: float) -> None:
    """Algorithmic (flops, bytes) of the NEXT C-ABI call -- minimal operand traffic, every tensor read
    or written once; only consumed while a timeline is being recorded."""
    global _work
    _work = (float(flops), float(nbytes))


def _call(name, *args):
    global _launches, _work
    _launches += 1
    if _timeline is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call(name, *args)
        b.record()
        _timeline.append((name, a, b, _work))
    else:
        _lib.call(name, *args)
    _work = (0.0, 0.0)
    if _SYNC_DEBUG:  # attribute asynchronous kernel faults to the entry point that caused them
        try:
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            raise _lib.XmodalError(f"{name} faulted on the device: {exc}") from exc


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor with unit inner stride, 16-B aligned base and row pitch % 4 == 0 (TMA-readable)."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.stride(1) != 1 or t.stride(0) % 4 != 0 or t.data_ptr() % 16 != 0 or t.stride(0) < t.size(1):
        k = t.size(1)
        kp = (k + 3) // 4 * 4
        buf = torch.zeros(t.size(0), kp, device=t.device, dtype=t.dtype)
        buf[:, :k] = t
        return buf[:, :k]
    return t


def pitch4(n: int) -> int:
    return (n + 3) // 4 * 4


def empty_pitched(shape: Tuple[int, ...], device) -> torch.Tensor:
    """(..., T) fp32 tensor whose last-dim pitch is a multiple of 4 elements (a view if T % 4)."""
    *lead, T = shape
    Tp = pitch4(T)
    buf = torch.empty(*lead, Tp, device=device, dtype=torch.float32)
    return buf if Tp == T else buf[..., :T]


def as_nwc(t: torch.Tensor) -> torch.Tensor:
    """(B, T, C) channels-last tensor with unit channel stride, row pitch % 4 == 0, rows densely
    stacked and a 16-B aligned base (what the TMA tensor maps need); copies only if it must."""
    B, T, C = t.shape
    ld = t.stride(2) == 1 and t.stride(1)
    ok = ld and ld % 4 == 0 and ld >= C and t.stride(0) == T * ld and t.data_ptr() % 16 == 0
    if ok:
        return t
    out = empty_pitched((B, T, C), t.device)
    out.copy_(t)
    return out


def to_nwc(x: torch.Tensor, round_out: bool = False, split3: bool = False) -> torch.Tensor:
    """(B, C, T) reference layout -> channels-last (B, T, C) (pitched), one transposing pass; split3: (B, T, 2C), the
    channel-stacked tf32 split [hi | lo] (operand of a 3-pass first conv, which reads it as [hi | lo | hi])."""
    return window_gather(x, x.shape[2], 1, channels_last=True, round_out=round_out, split3=split3)


# ------------------------------------------------------------------ linear
def linear_fwd(x, w, bias=None, act=None, round_out=False, splits: int = 0, fp32_accum=False):
    _chk(x, w, bias)
    x, w = _rowmajor(x), _rowmajor(w)
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    if splits <= 0:  # split K only when the output grid alone cannot fill the GPU
        tiles = ((M + 127) // 128) * max(1, (N + 255) // 256)
        splits = 1 if tiles >= 74 or K < 2048 else min(8, max(1, 148 // tiles), K // 512)
    ws = torch.empty(splits * M * N, device=x.device, dtype=torch.float32) if splits > 1 else None
    _w(2.0 * M * N * K, 4.0 * (M * K + N * K + M * N))
    _call("xm_linear_fwd_f32", _p(x), _p(w), _p(bias), _p(y), M, N, K, x.stride(0), w.stride(0), y.stride(0),
          act_code(act), int(round_out) | (2 if fp32_accum else 0), splits, _p(ws), _stream())
    return y


def linear_dgrad(dy, w, round_out=False):
    _chk(dy, w)
    dy, w = _rowmajor(dy), _rowmajor(w)
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
    _w(2.0 * M * N * K, 4.0 * (M * N + N * K + M * K))
    _call("xm_linear_dgrad_f32", _p(dy), _p(w), _p(dx), M, N, K, dy.stride(0), w.stride(0), dx.stride(0),
          int(round_out), _stream())
    return dx


def linear_wgrad(dy, x, need_bias=True, splits: int = 0):
    """dw (N, K) = dy^T x, db (N) = column sums of dy.  When the layer is wide on the output side (N > K,
    e.g. the packed q/k/v projection 128 -> 384 or the FFN up-projection 128 -> 512) the contraction runs as
    dw^T = x^T dy: the narrow side becomes the 128-row MMA tile and the wide side the N = 256 MMA width, which
    halves the shared-memory operand traffic per FLOP of the tensor core (small-N MMAs are smem-bound)."""
    _chk(dy, x)
    M, N = dy.shape
    K = x.shape[1]
    if N > K and N >= 256 and splits <= 0:
        dwt, _ = _linear_wgrad_nk(x, dy, False, 0)  # (K, N)
        return dwt.t().contiguous(), (colsum(_rowmajor(dy)) if need_bias else None)
    return _linear_wgrad_nk(dy, x, need_bias, splits)


def _linear_wgrad_nk(dy, x, need_bias=True, splits: int = 0):
    dy, x = _rowmajor(dy), _rowmajor(x)
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty(N, K, device=dy.device, dtype=torch.float32)
    db = torch.empty(N, device=dy.device, dtype=torch.float32) if need_bias else None
    if splits <= 0:
        tiles = ((N + 127) // 128) * max(1, (K + 255) // 256)
        splits = 1 if tiles >= 74 or M < 1024 else min(max(1, 148 // tiles), M // 256)
    ws = torch.empty(splits * N * K, device=dy.device, dtype=torch.float32) if splits > 1 else None
    dbws = _colsum_ws(M, N, dy.device) if need_bias else None
    _w(2.0 * M * N * K, 4.0 * (M * N + M * K + N * K))
    _call("xm_linear_wgrad_f32", _p(dy), _p(x), _p(dw), _p(db), M, N, K, dy.stride(0), x.stride(0), dw.stride(0),
          splits, _p(ws), _p(dbws), _stream())
    return dw, db


# ------------------------------------------------------------------ conv1d (channels-last activations)
def conv1d_pack_weight(w):
    """(Cout, Cin, taps) -> tf32-rounded (taps, Cout, ldk) and (taps, Cin, ldt) operand copies."""
    _chk(w)
    w = w.contiguous()
    Cout, Cin, taps = w.shape
    ldk, ldt = pitch4(Cin), pitch4(Cout)
    wk = torch.empty(taps, Cout, ldk, device=w.device, dtype=torch.float32)
    wt = torch.empty(taps, Cin, ldt, device=w.device, dtype=torch.float32)
    _call("xm_conv1d_pack_weight_f32", _p(w), Cout, Cin, taps, _p(wk), ldk, _p(wt), ldt, _stream())
    return wk, wt


def conv1d_fwd(x, wk, bias, Cout, round_out=False, out=None, stats=False, wrap_cin=0):
    """x (B, T, Cin) channels-last -> y (B, T, Cout).  `out` may be a channel slice of a wider
    (B, T, Ctot) buffer (free concat of parallel branches).  stats: -> (y, part) with part (rows, Cout, 2) fp64 = the
    BatchNorm partial statistics of y accumulated in the conv epilogue (what bn_partial_stats(y) returns, other split).
    wrap_cin > x.shape[2]: the contraction walks wrap_cin channels, x stores the first x.shape[2] of them and the rest
    wrap around ([hi | lo] read as [hi | lo | hi] by a 3-pass conv)."""
    _chk(x, wk, bias)
    x = as_nwc(x)
    B, T, Cx = x.shape
    Cin = int(wrap_cin) if wrap_cin else Cx
    taps, _, ldk = wk.shape
    y = empty_pitched((B, T, Cout), x.device) if out is None else out
    _w(2.0 * B * T * Cin * Cout * taps, 4.0 * (B * T * (Cx + Cout) + taps * Cin * Cout))
    st = _stream()

    def attempt(xx, want_part, x_cols):
        """One launch of the statistics / wrap entry point; False when the shape is outside it (XM_ERR_UNSUPPORTED)."""
        try:
            _call("xm_conv1d_fwd_stats_f32", _p(xx), _p(wk), _p(bias), _p(y), _p(want_part), B, Cin, Cout, T, taps, xx.stride(1), ldk,
                  y.stride(1), int(round_out), x_cols, st)
            return True
        except _lib.XmodalError as exc:
            if getattr(exc, "status", 0) != -2:
                raise
            return False

    wrapped = Cin != Cx
    if stats:
        part = torch.empty(_lib.lib().xm_conv1d_fwd_stat_rows(), Cout, 2, device=x.device, dtype=torch.float64)
        if attempt(x, part, Cx if wrapped else 0):
            return y, part
    if wrapped and not attempt(x, None, Cx):  # the wrap needs the halo producer: else materialise the wrapped blocks
        x = torch.cat([x, x[:, :, :Cin - Cx]], dim=2)
        wrapped = False
    if not wrapped:
        _call("xm_conv1d_fwd_f32", _p(x), _p(wk), _p(bias), _p(y), B, Cin, Cout, T, taps, x.stride(1), ldk, y.stride(1),
              int(round_out), st)
    return (y, bn_partial_stats(y)) if stats else y


def conv1d_fwd_precise(x, w, bias, stats=False):
    """fp32-accurate conv forward on the tf32 tensor cores: x (B, T, Cin) channels-last and NOT rounded, w (Cout, Cin,
    taps) in the reference layout.  x = xh + xl, w = wh + wl, y = xh wh + xl wh + xh wl as ONE conv over 3 Cin
    stacked channels [xh | xl | xh] x [wh | wh | wl]; x may also be the split [xh | xl] already, (B, T, 2 Cin).
    -> (y (B, T, Cout), xh view (B, T, Cin): the tf32-rounded input, what a single-pass weight gradient reads)."""
    _chk(x, w, bias)
    x = as_nwc(x)
    B, T, _ = x.shape
    Cout, Cin, taps = w.shape
    wrap = 0
    if x.shape[2] == 2 * Cin and Cin != 0:  # the producer already wrote the split [hi | lo] (window gather / BatchNorm block)
        x3, wrap = x, 3 * Cin                # read as [hi | lo | hi]: the third channel block wraps onto the first
        if (2 * Cin) % 32 or taps == 1:      # (the wrap needs whole 32-channel blocks and the halo producer)
            x3, wrap = torch.cat([x, x[:, :, :Cin]], dim=2), 0
    else:
        if x.stride(1) != Cin:
            x = x.contiguous()
        x3 = split3(x.view(B * T, Cin), 0, 1).view(B, T, 3 * Cin)
    w3 = split3(w.reshape(Cout, Cin * taps), 1, 1).view(Cout, 3 * Cin, taps)
    wk3, _ = conv1d_pack_weight(w3)
    if stats:
        y, part = conv1d_fwd(x3, wk3, bias, Cout, stats=True, wrap_cin=wrap)
        return y, x3[:, :, :Cin], part
    return conv1d_fwd(x3, wk3, bias, Cout, wrap_cin=wrap), x3[:, :, :Cin]


def conv1d_dgrad(dy, wt, Cin, round_out=False, out=None):
    """`out`: write dx into this (B, T, Cin) view of a wider buffer (row pitch = out.stride(1))."""
    _chk(dy, wt)
    dy = as_nwc(dy)
    B, T, Cout = dy.shape
    taps, _, ldt = wt.shape
    dx = empty_pitched((B, T, Cin), dy.device) if out is None else out
    _w(2.0 * B * T * Cin * Cout * taps, 4.0 * (B * T * (Cin + Cout) + taps * Cin * Cout))
    _call("xm_conv1d_dgrad_f32", _p(dy), _p(wt), _p(dx), B, Cin, Cout, T, taps, dy.stride(1), ldt, dx.stride(1),
          int(round_out), _stream())
    return dx


def conv1d_wgrad(dy, x, taps, need_bias=True):
    """dy (B, T, Cout), x (B, T, Cin) channels-last -> dw (Cout, Cin, taps) [reference layout], db."""
    _chk(dy, x)
    dy, x = as_nwc(dy), as_nwc(x)
    B, T, Cout = dy.shape
    Cin = x.shape[2]
    n_ws = _lib.lib().xm_conv1d_wgrad_workspace(B, Cin, Cout, taps)
    ws = torch.empty(n_ws, device=dy.device, dtype=torch.float32)
    dw = torch.empty(Cout, Cin, taps, device=dy.device, dtype=torch.float32)
    db = torch.empty(Cout, device=dy.device, dtype=torch.float32) if need_bias else None
    _w(2.0 * B * T * Cin * Cout * taps, 4.0 * (B * T * (Cin + Cout) + taps * Cin * Cout))
    dbws = _colsum_ws(B * T, Cout, dy.device) if need_bias else None
    _call("xm_conv1d_wgrad_f32", _p(dy), _p(x), _p(dw), _p(db), B, Cin, Cout, T, taps, dy.stride(1), x.stride(1),
          _p(ws), _p(dbws), _stream())
    return dw, db


# ------------------------------------------------------------------ batch norm + act (+pool, +dropout)
def _bn_dims(y):
    """-> (B, T, C, ld) for channels-last (B, T, C) or (B, C) inputs."""
    if y.dim() == 2:
        assert y.stride(1) == 1
        return y.shape[0], 1, y.shape[1], y.stride(0)
    B, T, C = y.shape
    assert y.stride(2) == 1 and y.stride(0) == T * y.stride(1)
    return B, T, C, y.stride(1)


def bn_partial_stats(y):
    """Per-split {sum, sumsq} doubles, shape (nsplit, C, 2)."""
    _chk(y)
    B, T, C, ld = _bn_dims(y)
    ns = _lib.lib().xm_bn_nsplit(B * T, C)
    part = torch.empty(ns, C, 2, device=y.device, dtype=torch.float64)
    _w(3.0 * B * T * C, 4.0 * B * T * C)
    _call("xm_bn_partial_stats_f32", _p(y), B * T, C, ld, _p(part), _stream())
    return part


def bn_finalize_stats(part, count, eps, running_mean=None, running_var=None, momentum=0.1):
    ns, C, _ = part.shape
    mean = torch.empty(C, device=part.device, dtype=torch.float32)
    invstd = torch.empty(C, device=part.device, dtype=torch.float32)
    _call("xm_bn_finalize_stats", _p(part), ns, C, float(count), float(eps), _p(mean), _p(invstd), _p(running_mean),
          _p(running_var), float(momentum), _stream())
    return mean, invstd


def bn_act_fwd(y, mean, invstd, gamma, beta, act, pool=0, drop_p=0.0, seed=0, drop_before_pool=False,
               round_out=False):
    _chk(y, mean, invstd, gamma, beta)
    B, T, C, ld = _bn_dims(y)
    split = int(round_out) == 2  # (B, T', 2C): channel-stacked tf32 split [hi | lo] for a following 3-pass conv
    if y.dim() == 2:
        out = torch.empty(B, 2 * C if split else C, device=y.device, dtype=torch.float32)
        ldo = out.stride(0)
    else:
        out = empty_pitched((B, T // 2 if pool == 2 else T, 2 * C if split else C), y.device)
        ldo = out.stride(1)
    _w(20.0 * B * T * C, 4.0 * (B * T * C + out.numel()))
    _call("xm_bn_act_fwd_f32", _p(y), _p(mean), _p(invstd), _p(gamma), _p(beta), _p(out), B, T, C, ld, ldo,
          act_code(act), pool, float(drop_p), int(seed), int(drop_before_pool), int(round_out), _stream())
    return out


def _ldo(dout):
    return dout.stride(0) if dout.dim() == 2 else dout.stride(1)


def bn_act_bwd_reduce(dout, y, mean, invstd, gamma, beta, act, pool=0, drop_p=0.0, seed=0, drop_before_pool=False):
    _chk(dout, y)
    B, T, C, ld = _bn_dims(y)
    ns = _lib.lib().xm_bn_nsplit(B * T, C)
    part = torch.empty(ns, C, 2, device=y.device, dtype=torch.float64)
    _w(30.0 * B * T * C, 4.0 * (B * T * C + dout.numel()))
    _call("xm_bn_act_bwd_reduce_f32", _p(dout), _p(y), _p(mean), _p(invstd), _p(gamma), _p(beta), B, T, C, ld,
          _ldo(dout), act_code(act), pool, float(drop_p), int(seed), int(drop_before_pool), _p(part), _stream())
    return part


def bn_bwd_finalize(part):
    ns, C, _ = part.shape
    dbeta = torch.empty(C, device=part.device, dtype=torch.float32)
    dgamma = torch.empty(C, device=part.device, dtype=torch.float32)
    _call("xm_bn_bwd_finalize", _p(part), ns, C, _p(dbeta), _p(dgamma), _stream())
    return dbeta, dgamma


def bn_act_bwd_apply(dout, y, mean, invstd, gamma, beta, dbeta, dgamma, count, act, pool=0, drop_p=0.0, seed=0,
                     drop_before_pool=False, round_out=False):
    B, T, C, ld = _bn_dims(y)
    dy = torch.empty(B, C, device=y.device, dtype=torch.float32) if y.dim() == 2 else empty_pitched((B, T, C), y.device)
    lddy = dy.stride(0) if y.dim() == 2 else dy.stride(1)
    assert lddy == ld, "dy is written with y's pitch"
    _w(30.0 * B * T * C, 4.0 * (2 * B * T * C + dout.numel()))
    _call("xm_bn_act_bwd_apply_f32", _p(dout), _p(y), _p(mean), _p(invstd), _p(gamma), _p(beta), _p(dbeta),
          _p(dgamma), float(count), _p(dy), B, T, C, ld, _ldo(dout), act_code(act), pool, float(drop_p), int(seed),
          int(drop_before_pool), int(round_out), _stream())
    return dy


def seqmean(x):
    """x (B, T, C) channels-last -> (B, C): AdaptiveAvgPool1d(1)."""
    _chk(x)
    B, T, C, ld = _bn_dims(x)
    out = torch.empty(B, C, device=x.device, dtype=torch.float32)
    _w(B * T * C, 4.0 * (B * T * C + B * C))
    _call("xm_seqmean_f32", _p(x), B, T, C, ld, _p(out), _stream())
    return out


def seqmean_bwd(dout, T):
    _chk(dout)
    dout = dout.contiguous()
    B, C = dout.shape
    dx = empty_pitched((B, T, C), dout.device)
    _w(B * T * C, 4.0 * (B * T * C + B * C))
    _call("xm_seqmean_bwd_f32", _p(dout), B, T, C, dx.stride(1), _p(dx), _stream())
    return dx


# ------------------------------------------------------------------ layer norm + act
def ln_act_fwd(x, gamma, beta, eps, act, drop_p=0.0, seed=0):
    _chk(x, gamma, beta)
    x = x.contiguous()
    M, D = x.shape
    out = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    _w(20.0 * M * D, 8.0 * M * D)
    _call("xm_ln_act_fwd_f32", _p(x), _p(gamma), _p(beta), _p(out), _p(mean), _p(rstd), M, D, float(eps),
          act_code(act), float(drop_p), int(seed), _stream())
    return out, mean, rstd


def ln_act_bwd(dout, x, gamma, beta, mean, rstd, act, drop_p=0.0, seed=0):
    _chk(dout, x)
    dout, x = dout.contiguous(), x.contiguous()
    M, D = x.shape
    nblk = _lib.lib().xm_ln_nblk(M)
    dx = torch.empty_like(x)
    dgp = torch.empty(nblk, D, device=x.device, dtype=torch.float32)
    dbp = torch.empty(nblk, D, device=x.device, dtype=torch.float32)
    _w(30.0 * M * D, 12.0 * M * D)
    _call("xm_ln_act_bwd_f32", _p(dout), _p(x), _p(gamma), _p(beta), _p(mean), _p(rstd), _p(dx), _p(dgp), _p(dbp), M,
          D, act_code(act), float(drop_p), int(seed), _stream())
    return dx, colsum(dgp), colsum(dbp)


# ------------------------------------------------------------------ activation + dropout
def act_fwd(x, act, drop_p=0.0, seed=0, round_out=False):
    _chk(x)
    x = x.contiguous()
    out = torch.empty_like(x)
    _w(10.0 * x.numel(), 8.0 * x.numel())
    _call("xm_act_fwd_f32", _p(x), _p(out), x.numel(), act_code(act), float(drop_p), int(seed), int(round_out), _stream())
    return out


def act_bwd(dout, x, act, drop_p=0.0, seed=0, round_out=False):
    _chk(dout, x)
    dout, x = dout.contiguous(), x.contiguous()
    dx = torch.empty_like(x)
    _w(10.0 * x.numel(), 12.0 * x.numel())
    _call("xm_act_bwd_f32", _p(dout), _p(x), _p(dx), x.numel(), act_code(act), float(drop_p), int(seed), int(round_out),
          _stream())
    return dx


def split3(x, which, axis):
    """3-way tf32 split of a 2-D tensor along `axis` ([hi|lo|hi] for which=0, [hi|hi|lo] for which=1)."""
    _chk(x)
    x = x.contiguous()
    R, C = x.shape
    out = torch.empty((3 * R, C) if axis == 0 else (R, 3 * C), device=x.device, dtype=torch.float32)
    _w(3.0 * R * C, 16.0 * R * C)
    _call("xm_split3_f32", _p(x), _p(out), R, C, int(which), int(axis), _stream())
    return out


def linear_fwd_precise(x, w, bias=None, act=None):
    """fp32-accurate y = x @ w^T + b: three tf32 passes fused into one GEMM over a tripled K."""
    return linear_fwd(split3(x, 0, 1), split3(w, 1, 1), bias, act=act, fp32_accum=True)


def linear_precise_prepare(x):
    """Row-stacked tf32 split [hi; hi; lo] of x (M, K): the operand of linear_fwd_prepared AND of the weight gradient."""
    if x.shape[1] % 4:  # TMA row pitch: zero columns up to a multiple of 4 (they contribute nothing)
        x = torch.nn.functional.pad(x, (0, 4 - x.shape[1] % 4))
    return split3(x, 1, 0)


def linear_fwd_prepared(x3, w, bias=None, act=None):
    """fp32-accurate y = x @ w^T + b from the row-stacked split of x (see linear_precise_prepare)."""
    _chk(x3, w, bias)
    M, K = x3.shape[0] // 3, x3.shape[1]
    N = w.shape[0]
    if w.shape[1] != K:
        w = torch.nn.functional.pad(w, (0, K - w.shape[1]))
    w3 = split3(w, 0, 0)
    y = torch.empty(M, N, device=x3.device, dtype=torch.float32)
    tiles = ((M + 127) // 128) * max(1, (N + 127) // 128)
    # split K so that the persistent grid fills about two waves of the 148 SMs (32 row tiles x 9 splits = 288)
    splits = 1 if tiles >= 74 or K < 2048 else min(16, max(1, round(296 / tiles)), K // 512)
    ws = torch.empty(splits * M * N, device=x3.device, dtype=torch.float32) if splits > 1 else None
    _w(6.0 * M * N * K, 4.0 * (3 * M * K + 3 * N * K + M * N))
    _call("xm_linear_fwd_stacked3_f32", _p(x3), _p(w3), _p(bias), _p(y), M, N, K, x3.stride(0), w3.stride(0), y.stride(0),
          act_code(act), 2, splits, _p(ws), _stream())
    return y


def linear_wgrad_prepared(dy, x3, need_bias=True, K=None):
    """dw (N, K) from the row-stacked split of x; K: the layer's true input width when x3 carries padding columns."""
    dw, _ = linear_wgrad(split3(dy, 0, 0), x3, need_bias=False)
    if K is not None and K != dw.shape[1]:
        dw = dw[:, :K].contiguous()
    return dw, (colsum(dy) if need_bias else None)


def linear_dgrad_precise(dy, w):
    return linear_dgrad(split3(dy, 0, 1), split3(w, 1, 0))


def linear_wgrad_precise(dy, x, need_bias=True):
    dw, _ = linear_wgrad(split3(dy, 0, 0), split3(x, 1, 0), need_bias=False)
    return dw, (colsum(dy) if need_bias else None)


def act_bwd_colsum(dout, x, act, drop_p=0.0, seed=0, round_out=False):
    """act_bwd on (M, C) rows + the column sums of the result (bias gradient of the preceding Linear)."""
    _chk(dout, x)
    dout, x = dout.contiguous(), x.contiguous()
    M, C = x.shape
    nblk = _lib.lib().xm_act_bwd_colsum_nblk(M, C)
    if nblk == 0:
        dx = act_bwd(dout, x, act, drop_p, seed, round_out)
        return dx, colsum(dx)
    dx = torch.empty_like(x)
    part = torch.empty(nblk, C, device=x.device, dtype=torch.float32)
    _w(10.0 * x.numel(), 12.0 * x.numel())
    _call("xm_act_bwd_colsum_f32", _p(dout), _p(x), _p(dx), M, C, act_code(act), float(drop_p), int(seed), int(round_out),
          _p(part), _stream())
    return dx, colsum(part)


def round_tf32(x, inplace=False):
    """Round to nearest tf32 (what the tensor cores would otherwise truncate to)."""
    _chk(x)
    x = x.contiguous()
    out = x if inplace else torch.empty_like(x)
    _w(x.numel(), 8.0 * x.numel())
    _call("xm_round_tf32_f32", _p(x), _p(out), x.numel(), _stream())
    return out


# ------------------------------------------------------------------ reductions
def _colsum_ws(M, N, device):
    ns = _lib.lib().xm_colsum_nsplit(M, N)
    return torch.empty(ns * N, device=device, dtype=torch.float32) if ns > 1 else None


def colsum(x):
    _chk(x)
    M, N = x.shape
    assert x.stride(1) == 1
    out = torch.empty(N, device=x.device, dtype=torch.float32)
    ws = _colsum_ws(M, N, x.device)
    _w(M * N, 4.0 * (M * N + N))
    _call("xm_colsum_f32", _p(x), M, N, x.stride(0), _p(out), _p(ws), _stream())
    return out


# ------------------------------------------------------------------ l2norm / similarity / infonce
def l2norm_fwd(x, eps=1e-12):
    _chk(x)
    x = x.contiguous()
    M, D = x.shape
    xn = torch.empty_like(x)
    inv = torch.empty(M, device=x.device, dtype=torch.float32)
    _w(3.0 * M * D, 8.0 * M * D)
    _call("xm_l2norm_fwd_f32", _p(x), _p(xn), _p(inv), M, D, float(eps), _stream())
    return xn, inv


def l2norm_split_fwd(x, which, eps=1e-12, xs_out=None):
    """-> (xn fp32 (M, D), xs (M, 3D) tf32 split [hi|lo|hi] (which=0) or [hi|hi|lo] (which=1), inv_norm).
    `xs_out`: write the split into this (M, 3D) buffer (e.g. NVLink-mapped symmetric memory)."""
    _chk(x)
    x = x.contiguous()
    M, D = x.shape
    xn = torch.empty_like(x)
    xs = torch.empty(M, 3 * D, device=x.device, dtype=torch.float32) if xs_out is None else xs_out
    inv = torch.empty(M, device=x.device, dtype=torch.float32)
    _w(6.0 * M * D, 20.0 * M * D)
    _call("xm_l2norm_split_fwd_f32", _p(x), _p(xn), _p(xs), _p(inv), M, D, float(eps), int(which), _stream())
    return xn, xs, inv


def l2norm_bwd(dxn, xn, inv):
    _chk(dxn, xn, inv)
    dxn = dxn.contiguous()
    M, D = xn.shape
    dx = torch.empty_like(xn)
    _w(4.0 * M * D, 12.0 * M * D)
    _call("xm_l2norm_bwd_f32", _p(dxn), _p(xn), _p(inv), _p(dx), M, D, _stream())
    return dx


def similarity(a, b, inv_tau):
    _chk(a, b)
    a, b = a.contiguous(), b.contiguous()
    Ml, D = a.shape
    Ng = b.shape[0]
    S = torch.empty(Ml, Ng, device=a.device, dtype=torch.float32)
    _w(2.0 * Ml * Ng * D, 4.0 * (Ml * D + Ng * D + Ml * Ng))
    _call("xm_similarity_f32", _p(a), _p(b), _p(S), Ml, Ng, D, float(inv_tau), _stream())
    return S


def infonce_lse(a, b, inv_tau, diag_off=0):
    _chk(a, b)
    a, b = a.contiguous(), b.contiguous()
    Ml, D = a.shape
    Ng = b.shape[0]
    tn = _lib.lib().xm_infonce_tile_n()
    ws = torch.empty(Ml * ((Ng + tn - 1) // tn), device=a.device, dtype=torch.float32)
    lse = torch.empty(Ml, device=a.device, dtype=torch.float32)
    diag = torch.empty(Ml, device=a.device, dtype=torch.float32)
    _w(2.0 * Ml * Ng * D, 4.0 * (Ml * D + Ng * D + 2 * Ml))
    _call("xm_infonce_lse_f32", _p(a), _p(b), _p(lse), _p(diag), Ml, Ng, D, float(inv_tau), int(diag_off), _p(ws),
          _stream())
    return lse, diag


def infonce_grad(a, b, lse_row, lse_col, inv_tau, diag_off, coef, round_out=False):
    _chk(a, b, lse_row, lse_col)
    a, b = a.contiguous(), b.contiguous()
    Ml, D = a.shape
    Ng = b.shape[0]
    G = torch.empty(Ml, Ng, device=a.device, dtype=torch.float32)
    _w(2.0 * Ml * Ng * D, 4.0 * (Ml * D + Ng * D + Ml * Ng))
    _call("xm_infonce_grad_f32", _p(a), _p(b), _p(lse_row), _p(lse_col), _p(G), Ml, Ng, D, float(inv_tau),
          int(diag_off), float(coef), int(round_out), _stream())
    return G


def _ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def peer_gather(peer_ptrs, rows_per_peer, cols, device):
    """(G*rows_per_peer, cols) local copy of the ranks' shards, read through their NVLink mappings."""
    G = len(peer_ptrs)
    out = torch.empty(G * rows_per_peer, cols, device=device, dtype=torch.float32)
    _w(0.0, 8.0 * out.numel())
    _call("xm_peer_gather_f32", _ptr_array(peer_ptrs), G, rows_per_peer * cols, _p(out), _stream())
    return out


def peer_allreduce_f64(x, data_dst, flag_dst, slot_data_ptr, slot_flags_ptr, row_stride, seq):
    """Sum of the ranks' fp64 vectors `x` through peer memory (xm_peer_allreduce_f64); pointers are raw device
    addresses into the symmetric buffers (functional._PeerReduce owns slots and sequence numbers)."""
    if not x.is_cuda or x.dtype != torch.float64:
        raise _lib.XmodalError(f"expected a CUDA float64 tensor, got {x.device} {x.dtype}")
    x = x.contiguous()
    out = torch.empty_like(x)
    _call("xm_peer_allreduce_f64", _p(x), _p(out), x.numel(), _ptr_array(data_dst), _ptr_array(flag_dst), len(data_dst),
          ctypes.c_void_p(int(slot_data_ptr)), ctypes.c_void_p(int(slot_flags_ptr)), int(row_stride), int(seq), _stream())
    return out


def infonce_lse_peers(a, peer_ptrs, rows_per_peer, inv_tau, diag_off=0):
    """Row logsumexp of a @ [b_0; b_1; ...]^T * inv_tau where shard r of the second operand is read IN PLACE
    from rank r's memory (peer_ptrs[r], NVLink-mapped): all-gather fused into the GEMM's TMA loads."""
    _chk(a)
    a = a.contiguous()
    Ml, D = a.shape
    G = len(peer_ptrs)
    Ng = G * rows_per_peer
    tn = _lib.lib().xm_infonce_tile_n()
    ws = torch.empty(Ml * ((Ng + tn - 1) // tn), device=a.device, dtype=torch.float32)
    lse = torch.empty(Ml, device=a.device, dtype=torch.float32)
    diag = torch.empty(Ml, device=a.device, dtype=torch.float32)
    _w(2.0 * Ml * Ng * D, 4.0 * (Ml * D + Ng * D + 2 * Ml))
    _call("xm_infonce_lse_peers_f32", _p(a), _ptr_array(peer_ptrs), G, rows_per_peer, _p(lse), _p(diag), Ml, D,
          float(inv_tau), int(diag_off), _p(ws), _stream())
    return lse, diag


def infonce_grad_peers(a, peer_ptrs, rows_per_peer, lse_row, lse_col, inv_tau, diag_off, coef, round_out=False):
    _chk(a, lse_row, lse_col)
    a = a.contiguous()
    Ml, D = a.shape
    G_ = len(peer_ptrs)
    Ng = G_ * rows_per_peer
    G = torch.empty(Ml, Ng, device=a.device, dtype=torch.float32)
    _w(2.0 * Ml * Ng * D, 4.0 * (Ml * D + Ng * D + Ml * Ng))
    _call("xm_infonce_grad_peers_f32", _p(a), _ptr_array(peer_ptrs), G_, rows_per_peer, _p(lse_row), _p(lse_col), _p(G), Ml,
          D, float(inv_tau), int(diag_off), float(coef), int(round_out), _stream())
    return G


def linear_dgrad_peers(dy, peer_ptrs, rows_per_peer, K, ldw):
    """dx (M, K) = dy (M, G*rows_per_peer) @ w, w row-sharded across the peers (first K columns of pitch-ldw rows)."""
    _chk(dy)
    dy = _rowmajor(dy)
    M, N = dy.shape
    assert N == len(peer_ptrs) * rows_per_peer
    dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
    _w(2.0 * M * N * K, 4.0 * (M * N + N * K + M * K))
    _call("xm_linear_dgrad_peers_f32", _p(dy), _ptr_array(peer_ptrs), len(peer_ptrs), rows_per_peer, _p(dx), M, K,
          dy.stride(0), ldw, dx.stride(0), 0, _stream())
    return dx


# ------------------------------------------------------------------ multi-head self-attention core
def attn_supported(L: int, dh: int) -> bool:
    """Shapes of the fused tcgen05 attention core (xm_attn_fused_*); everything else: attn_general_*."""
    return dh == 32 and 0 < L <= 512


def attn_fused_fwd(qkv, nhead, scale, drop_p=0.0, seed=0, round_out=True):
    """Fused attention core: qkv (B, L, 3*H*dh) -> (out (B, L, H*dh), lse (B*H, L)); probabilities stay on chip."""
    _chk(qkv)
    qkv = qkv.contiguous()
    B, L, E = qkv.shape
    d = E // 3
    dh = d // nhead
    out = torch.empty(B, L, d, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(B * nhead, L, device=qkv.device, dtype=torch.float32)
    _w(4.0 * B * nhead * L * L * dh, 4.0 * (qkv.numel() + out.numel()))
    _call("xm_attn_fused_fwd_f32", _p(qkv), _p(out), _p(lse), B, L, nhead, dh, float(scale), float(drop_p), int(seed),
          int(round_out), _stream())
    return out, lse


def attn_fused_bwd(dout, qkv, out, lse, nhead, scale, drop_p=0.0, seed=0, round_out=False, need_bias=False):
    """-> dqkv, or (dqkv, dbias (3d)) with need_bias: the column sums of dqkv (the in-projection's bias gradient)
    accumulated inside the kernels."""
    _chk(dout, qkv, out, lse)
    dout, out = dout.contiguous(), out.contiguous()
    B, L, E = qkv.shape
    dh = E // 3 // nhead
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    part = torch.empty(_lib.lib().xm_attn_fused_bwd_nblk(), E, device=qkv.device, dtype=torch.float32) if need_bias else None
    _w(14.0 * B * nhead * L * L * dh, 4.0 * (3 * qkv.numel() + 3 * dout.numel()))
    _call("xm_attn_fused_bwd_f32", _p(dout), _p(qkv), _p(out), _p(lse), _p(dqkv), _p(delta), _p(part), B, L, nhead, dh,
          float(scale), float(drop_p), int(seed), int(round_out), _stream())
    return (dqkv, colsum(part)) if need_bias else dqkv


def attn_general_supported(L: int, dh: int) -> bool:
    return bool(_lib.lib().xm_attn_general_supported(int(L), int(dh)))


def attn_general_fwd(qkv, nhead, scale, mask=None, drop_p=0.0, seed=0, round_out=False):
    """Shape-general attention core (any head dim <= 256, optional additive mask (L, L) or (B*H, L, L))."""
    _chk(qkv, mask)
    qkv = qkv.contiguous()
    B, L, E = qkv.shape
    d = E // 3
    dh = d // nhead
    if mask is not None:
        mask = mask.contiguous()
        if tuple(mask.shape) not in ((L, L), (B * nhead, L, L)):
            raise _lib.XmodalError(f"attention mask {tuple(mask.shape)}: expected ({L}, {L}) or ({B * nhead}, {L}, {L})")
    out = torch.empty(B, L, d, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(B * nhead, L, device=qkv.device, dtype=torch.float32)
    _w(4.0 * B * nhead * L * L * dh, 4.0 * (qkv.numel() + out.numel()))
    _call("xm_attn_general_fwd_f32", _p(qkv), _p(mask), int(mask is not None and mask.dim() == 3), _p(out), _p(lse), B, L, nhead,
          dh, float(scale), float(drop_p), int(seed), int(round_out), _stream())
    return out, lse


def attn_general_bwd(dout, qkv, lse, nhead, scale, mask=None, drop_p=0.0, seed=0, round_out=False):
    _chk(dout, qkv, lse, mask)
    dout, qkv = dout.contiguous(), qkv.contiguous()
    B, L, E = qkv.shape
    dh = E // 3 // nhead
    mask = None if mask is None else mask.contiguous()
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    _w(10.0 * B * nhead * L * L * dh, 4.0 * (2 * qkv.numel() + dout.numel()))
    _call("xm_attn_general_bwd_f32", _p(dout), _p(qkv), _p(mask), int(mask is not None and mask.dim() == 3), _p(lse), _p(dqkv),
          _p(delta), B, L, nhead, dh, float(scale), float(drop_p), int(seed), int(round_out), _stream())
    return dqkv


def attn_fused_mask(B, L, nhead, drop_p, seed, device="cuda"):
    """The keep mask (B*H, L, L) the fused attention kernels generate for (drop_p, seed)."""
    mask = torch.empty(B * nhead, L, L, device=device, dtype=torch.uint8)
    _call("xm_attn_fused_mask_u8", _p(mask), B, L, nhead, float(drop_p), int(seed), _stream())
    return mask


# ------------------------------------------------------------------ fused feed-forward branch
def ffn_fused_supported(D: int, hidden: int, act) -> bool:
    return bool(_lib.lib().xm_ffn_fused_supported(int(D), int(hidden), act_code(act)))


def ffn_fused_fwd(x, w1, b1, w2, b2, act, drop_p=0.0, seed=0):
    """y = tf32(Dropout(act(x w1^T + b1))) w2^T + b2 with the hidden activations kept on chip; x, w1, w2 tf32-rounded."""
    _chk(x, w1, b1, w2, b2)
    x, w1, w2 = x.contiguous(), w1.contiguous(), w2.contiguous()
    M, D = x.shape
    H = w1.shape[0]
    y = torch.empty(M, D, device=x.device, dtype=torch.float32)
    _w(4.0 * M * D * H, 4.0 * (2 * M * D + 2 * D * H))
    _call("xm_ffn_fused_fwd_f32", _p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), M, D, H, act_code(act), float(drop_p), int(seed),
          _stream())
    return y


def ffn_fused_dgrad(x, dy, w1, b1, w2t, w1t, act, drop_p=0.0, seed=0):
    """-> (a (M, H), dh (M, H), dx (M, D), db1 (H)): see xm_ffn_fused_dgrad_f32; w2t = w2^T, w1t = w1^T (tf32 copies)."""
    _chk(x, dy, w1, b1, w2t, w1t)
    x, dy, w1, w2t, w1t = x.contiguous(), dy.contiguous(), w1.contiguous(), w2t.contiguous(), w1t.contiguous()
    M, D = x.shape
    H = w1.shape[0]
    a = torch.empty(M, H, device=x.device, dtype=torch.float32)
    dh = torch.empty(M, H, device=x.device, dtype=torch.float32)
    dx = torch.empty(M, D, device=x.device, dtype=torch.float32)
    part = torch.empty(_lib.lib().xm_ffn_fused_nblk(M), H, device=x.device, dtype=torch.float32)
    _w(6.0 * M * D * H, 4.0 * (3 * M * D + 2 * M * H + 3 * D * H))
    _call("xm_ffn_fused_dgrad_f32", _p(x), _p(dy), _p(w1), _p(b1), _p(w2t), _p(w1t), _p(a), _p(dh), _p(dx), _p(part), M, D, H,
          act_code(act), float(drop_p), int(seed), _stream())
    return a, dh, dx, colsum(part)


def ffn_fused_mask(M, hidden, drop_p, seed, device="cuda"):
    mask = torch.empty(M, hidden, device=device, dtype=torch.uint8)
    _call("xm_ffn_fused_mask_u8", _p(mask), M, hidden, float(drop_p), int(seed), _stream())
    return mask


# ------------------------------------------------------------------ residual stream (transformer block)
def resid_ln_supported(D: int) -> bool:
    return D % 128 == 0 and 128 <= D <= 512


def resid_ln_fwd(x, a, gamma, beta, eps, drop_p=0.0, seed=0, pe=None, L=0):
    """s = x + Dropout(a)  |  Dropout(x + pe[row % L]);  h = tf32(LayerNorm(s)).  -> (s, h, mean, rstd); s is x
    itself when there is neither a branch nor a positional table."""
    _chk(x, a, gamma, beta, pe)
    x = x.contiguous()
    M, D = x.shape
    a = None if a is None else a.contiguous()
    s = torch.empty_like(x) if (a is not None or pe is not None) else None
    h = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    _w(12.0 * M * D, 4.0 * M * D * (2 + (a is not None) + (s is not None)))
    _call("xm_resid_ln_fwd_f32", _p(x), _p(a), _p(pe), int(L), _p(gamma), _p(beta), _p(s), _p(h), _p(mean), _p(rstd), M, D,
          float(eps), float(drop_p), int(seed), _stream())
    return (x if s is None else s), h, mean, rstd


def resid_ln_bwd(dh, dres, s, gamma, mean, rstd, drop_p=0.0, seed=0, need_da=True):
    """-> (dx, da | None, dgamma, dbeta, dabias | None); dabias = column sums of da (the bias gradient of the
    Linear that produced the branch), accumulated in the same pass."""
    _chk(dh, dres, s)
    dh, s = dh.contiguous(), s.contiguous()
    dres = None if dres is None else dres.contiguous()
    M, D = s.shape
    nblk = _lib.lib().xm_resid_ln_nblk(M)
    dx = torch.empty_like(s)
    da = torch.empty_like(s) if need_da else None
    parts = torch.empty(3 if need_da else 2, nblk, D, device=s.device, dtype=torch.float32)
    _w(20.0 * M * D, 4.0 * M * D * (3 + (dres is not None) + need_da))
    _call("xm_resid_ln_bwd_f32", _p(dh), _p(dres), _p(s), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(da), _p(parts[0]),
          _p(parts[1]), _p(parts[2]) if need_da else None, M, D, float(drop_p), int(seed), _stream())
    sums = colsum(parts.transpose(0, 1).reshape(nblk, -1))  # one reduction for all partial vectors
    return dx, da, sums[:D], sums[D:2 * D], (sums[2 * D:] if need_da else None)


def resid_seqmean_fwd(x, a, drop_p=0.0, seed=0):
    """x, a (B, T, D) -> (B, D) = mean_t (x + Dropout(a))."""
    _chk(x, a)
    x = x.contiguous()
    a = None if a is None else a.contiguous()
    B, T, D = x.shape
    out = torch.empty(B, D, device=x.device, dtype=torch.float32)
    _w(2.0 * x.numel(), 4.0 * x.numel() * (1 + (a is not None)))
    _call("xm_resid_seqmean_fwd_f32", _p(x), _p(a), B, T, D, _p(out), float(drop_p), int(seed), _stream())
    return out


def resid_seqmean_bwd(dout, T, drop_p=0.0, seed=0, need_dx=True, need_da=True):
    _chk(dout)
    dout = dout.contiguous()
    B, D = dout.shape
    dx = torch.empty(B, T, D, device=dout.device, dtype=torch.float32) if need_dx else None
    da = torch.empty(B, T, D, device=dout.device, dtype=torch.float32) if need_da else None
    _w(2.0 * B * T * D, 4.0 * B * T * D * (need_dx + need_da))
    _call("xm_resid_seqmean_bwd_f32", _p(dout), B, T, D, _p(dx), _p(da), float(drop_p), int(seed), _stream())
    return dx, da


def infonce_dgrad(G, f3, which_f):
    """dx = G @ f_n, fp32-accurate: G (Ml, Ng) fp32 is split 3-way complementary to f3 (Ng, 3D), the l2norm split
    of the unit vectors made with `which_f`."""
    _chk(G, f3)
    Ml, Ng = G.shape
    D = f3.shape[1] // 3
    if Ng % 4:  # TMA row pitch: pad the contraction with zero columns / rows
        pad = 4 - Ng % 4
        G = torch.nn.functional.pad(G, (0, pad))
        f3 = torch.nn.functional.pad(f3, (0, 0, 0, pad))
        Ng += pad
    g3 = split3(G, 1 - which_f, 1)
    f3 = f3.contiguous()
    dx = torch.empty(Ml, D, device=G.device, dtype=torch.float32)
    _w(6.0 * Ml * Ng * D, 4.0 * (3 * Ml * Ng + 3 * Ng * D + Ml * D))
    _call("xm_infonce_dgrad_f32", _p(g3), _p(f3), _p(dx), Ml, Ng, D, _stream())
    return dx


def infonce_bwd_fused_supported(Ml, Ng, D, diag_off):
    return bool(_lib.lib().xm_infonce_bwd_fused_supported(int(Ml), int(Ng), int(D), int(diag_off)))


def infonce_lse_fused(e3, f3, e3_all, f3_all, inv_tau, diag_off=0):
    """-> (lse_ef, lse_fe, diag): row logsumexps of (my e x all f) and (my f x all e) scores and the positives S_ii, one
    launch, scores never written (xm_infonce_lse_fused_f32; shapes as infonce_bwd_fused_supported)."""
    _chk(e3, f3, e3_all, f3_all)
    e3, f3, e3_all, f3_all = e3.contiguous(), f3.contiguous(), e3_all.contiguous(), f3_all.contiguous()
    Ml, D = e3.shape[0], e3.shape[1] // 3
    Ng = e3_all.shape[0]
    lse_ef = torch.empty(Ml, device=e3.device, dtype=torch.float32)
    lse_fe = torch.empty(Ml, device=e3.device, dtype=torch.float32)
    diag = torch.empty(Ml, device=e3.device, dtype=torch.float32)
    ws = torch.empty(int(_lib.lib().xm_infonce_lse_fused_workspace(Ml, Ng)), device=e3.device, dtype=torch.float32)
    _w(2.0 * 6.0 * Ml * Ng * D, 4.0 * (2 * 3 * Ml * D + 2 * 3 * Ng * D + 3 * Ml))
    _call("xm_infonce_lse_fused_f32", _p(e3), _p(f3), _p(e3_all), _p(f3_all), _p(lse_ef), _p(lse_fe), _p(diag), Ml, Ng, D,
          float(inv_tau), int(diag_off), _p(ws), _stream())
    return lse_ef, lse_fe, diag


def infonce_bwd_fused(e3, f3, e3_all, f3_all, lse_ef, lse_fe, lse_ef_all, lse_fe_all, inv_tau, diag_off, coef, precise=True):
    """-> (de, df) (Ml, D): the InfoNCE gradients with respect to this rank's unit embeddings, both softmax-gradient
    blocks formed and contracted on chip (xm_infonce_bwd_fused_f32).  e3 / f3: local l2norm splits (which 0 / 1),
    e3_all / f3_all: the global batch's."""
    _chk(e3, f3, e3_all, f3_all, lse_ef, lse_fe, lse_ef_all, lse_fe_all)
    e3, f3, e3_all, f3_all = e3.contiguous(), f3.contiguous(), e3_all.contiguous(), f3_all.contiguous()
    Ml, D = e3.shape[0], e3.shape[1] // 3
    Ng = e3_all.shape[0]
    de = torch.empty(Ml, D, device=e3.device, dtype=torch.float32)
    df = torch.empty(Ml, D, device=e3.device, dtype=torch.float32)
    ws = torch.empty(int(_lib.lib().xm_infonce_bwd_fused_workspace(Ml, Ng, D)), device=e3.device, dtype=torch.float32)
    passes = 6.0 if precise else 4.0  # 3 score passes + 3 / 1 contraction passes, two directions
    _w(2.0 * passes * 2.0 * Ml * Ng * D, 4.0 * (2 * 3 * Ml * D + 2 * 3 * Ng * D + 2 * 2 * Ng * D + 2 * Ml * D))
    _call("xm_infonce_bwd_fused_f32", _p(e3), _p(f3), _p(e3_all), _p(f3_all), _p(lse_ef.contiguous()), _p(lse_fe.contiguous()),
          _p(lse_ef_all.contiguous()), _p(lse_fe_all.contiguous()), _p(de), _p(df), Ml, Ng, D, float(inv_tau), int(diag_off),
          float(coef), int(bool(precise)), _p(ws), _stream())
    return de, df


# ------------------------------------------------------------------ preprocessing
def window_index(n_rec, n_samples, win, hop, rec_labels=None, rec_subjects=None, device="cuda"):
    n_win = (n_samples - win) // hop + 1
    tot = n_rec * n_win
    mk = lambda: torch.empty(tot, device=device, dtype=torch.int64)
    starts, rec_ids = mk(), mk()
    labels = mk() if rec_labels is not None else None
    subjects = mk() if rec_subjects is not None else None
    _call("xm_window_index_i64", n_rec, n_samples, win, hop, _p(rec_labels), _p(rec_subjects), _p(starts),
          _p(rec_ids), _p(labels), _p(subjects), _stream())
    return starts, rec_ids, labels, subjects


def window_gather(rec, win, hop, channels_last=False, round_out=False, split3=False):
    """rec (R, C, n) -> (R*n_win, C, win) [reference layout] or (R*n_win, win, C) [channels-last]; with split3
    (channels-last only) (R*n_win, win, 3C): the channel-stacked tf32 split [hi | lo | hi] of every window."""
    _chk(rec)
    rec = rec.contiguous()
    R, C, n = rec.shape
    n_win = (n - win) // hop + 1
    if split3:
        if not channels_last:
            raise _lib.XmodalError("split3 needs the channels-last layout")
        out = empty_pitched((R * n_win, win, 2 * C), rec.device)
    else:
        out = empty_pitched((R * n_win, win, C) if channels_last else (R * n_win, C, win), rec.device)
    _w(0.0, 4.0 * out.shape[0] * C * win * (3 if split3 else 2))
    _call("xm_window_gather_f32", _p(rec), R, C, n, win, hop, _p(out), out.stride(1), int(channels_last),
          2 if split3 else int(round_out), _stream())
    return out


def bandpower(rec, win, hop, nfft, fs, taper, taper_sumsq, band_bins, total_bins=None, path="auto"):
    """Band powers of every window.  total_bins: the host's sum of the band widths in `band_bins` -- with it the
    tensor-core DFT kernel can be chosen (few bins, >= 64 channels); path: "auto" | "fft" | "dft"."""
    _chk(rec, taper)
    rec = rec.contiguous()
    R, C, n = rec.shape
    n_win = (n - win) // hop + 1
    nb = band_bins.numel() // 2
    power = torch.empty(R * n_win, C, nb, device=rec.device, dtype=torch.float32)
    dft_ok = total_bins is not None and bool(_lib.lib().xm_bandpower_dft_supported(C, n, win, hop, nfft, nb, int(total_bins)))
    if path == "dft" and not dft_ok:
        raise ValueError("shape not eligible for the DFT band-power kernel")
    if path == "dft" or (path == "auto" and dft_ok and C >= 64):
        ws = torch.empty(int(_lib.lib().xm_bandpower_dft_workspace_floats(win)), device=rec.device, dtype=torch.float32)
        _w(R * n_win * C * 6.0 * 64 * win, 4.0 * (R * n_win * C * win + power.numel()))
        _call("xm_bandpower_dft_f32", _p(rec), R, C, n, win, hop, nfft, float(fs), _p(taper), float(taper_sumsq),
              _p(band_bins), nb, int(total_bins), _p(ws), _p(power), _stream())
        return power
    _w(R * n_win * C * 2.5 * nfft * max(1, nfft.bit_length() - 1), 4.0 * (R * n_win * C * win + power.numel()))
    _call("xm_bandpower_f32", _p(rec), R, C, n, win, hop, nfft, float(fs), _p(taper), float(taper_sumsq),
          _p(band_bins), nb, _p(power), _stream())
    return power


def zscore(x, eps=1e-8):
    """Per-item (dim 0) global z-score, population std."""
    _chk(x)
    x = x.contiguous()
    n = x.shape[0]
    out = torch.empty_like(x)
    _w(4.0 * x.numel(), 8.0 * x.numel())
    _call("xm_zscore_f32", _p(x), n, x.numel() // n, float(eps), _p(out), _stream())
    return out


def roi_meanstd(x):
    _chk(x)
    x = x.contiguous()
    B, TR, ROI = x.shape
    out = torch.empty(B, 2 * ROI, device=x.device, dtype=torch.float32)
    _w(3.0 * x.numel(), 4.0 * (x.numel() + out.numel()))
    _call("xm_roi_meanstd_f32", _p(x), B, TR, ROI, _p(out), _stream())
    return out


def roi_corrcoef(x, prepared=False):
    """x (B, TR, ROI) -> (B, ROI*ROI): flattened per-sample Pearson correlation matrix of the ROI columns over TR.
    prepared: (3B, ROI*ROI), the row-stacked tf32 split of that matrix (what linear_precise_prepare makes of it)."""
    _chk(x)
    x = x.contiguous()
    B, TR, ROI = x.shape
    if prepared and (ROI * ROI) % 4:
        raise _lib.XmodalError("prepared connectivity needs ROI*ROI % 4 == 0")
    out = torch.empty(3 * B if prepared else B, ROI * ROI, device=x.device, dtype=torch.float32)
    _w(2.0 * B * TR * ROI * ROI, 4.0 * (x.numel() + out.numel()))
    _call("xm_roi_corrcoef_f32", _p(x), B, TR, ROI, _p(out), int(prepared), _stream())
    return out


# ------------------------------------------------------------------ optimizer step over one flat bucket
def cross2_attn_fwd(q, kv, nhead, drop_p=0.0, seed=0):
    """Bridge head: one query per sample over two tokens.  q (B, d), kv (2B, 2d) -> (out (B, d), att (B, H, 2))."""
    _chk(q, kv)
    q, kv = q.contiguous(), kv.contiguous()
    B, d = q.shape
    out = torch.empty(B, d, device=q.device, dtype=torch.float32)
    att = torch.empty(B, nhead, 2, device=q.device, dtype=torch.float32)
    _w(8.0 * B * d, 4.0 * (6 * B * d))
    _call("xm_cross2_attn_fwd_f32", _p(q), _p(kv), _p(out), _p(att), B, nhead, d // nhead, float(drop_p), int(seed), _stream())
    return out, att


def cross2_attn_bwd(dout, q, kv, nhead, drop_p=0.0, seed=0):
    """-> (dq (B, d), dkv (2B, 2d))"""
    _chk(dout, q, kv)
    dout, q, kv = dout.contiguous(), q.contiguous(), kv.contiguous()
    B, d = q.shape
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    _w(16.0 * B * d, 4.0 * (10 * B * d))
    _call("xm_cross2_attn_bwd_f32", _p(dout), _p(q), _p(kv), _p(dq), _p(dkv), B, nhead, d // nhead, float(drop_p), int(seed),
          _stream())
    return dq, dkv


def gather_flat_(dst, tensors, offsets):
    """dst[offsets[t] : offsets[t] + tensors[t].numel()] = tensors[t] (flattened), all tensors in one launch per 96:
    the per-parameter gradients autograd produced, gathered into the flat bucket (xm_gather_flat_f32)."""
    if not tensors:
        return dst
    srcs = []
    for t in tensors:
        if not t.is_cuda or t.dtype != torch.float32:
            raise _lib.XmodalError(f"expected CUDA float32 gradients, got {t.device} {t.dtype}")
        srcs.append(t if t.is_contiguous() else t.contiguous())
    n = len(srcs)
    off = (ctypes.c_int64 * n)(*[int(o) for o in offsets])
    num = (ctypes.c_int64 * n)(*[int(t.numel()) for t in srcs])
    _w(0.0, 8.0 * sum(t.numel() for t in srcs))
    _call("xm_gather_flat_f32", _ptr_array([t.data_ptr() for t in srcs]), off, num, n, _p(dst), _stream())
    return dst


def clip_adamw_(p, g, m, v, step, lr, weight_decay, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8):
    """In place on flat fp32 buffers: clip_grad_norm_(max_norm) + torch.optim.AdamW update for 1-based `step`.
    -> the pre-clip total gradient norm (1-element device tensor)."""
    _chk(p, g, m, v)
    n = p.numel()
    nblk = _lib.lib().xm_sumsq_nblk(n)
    part = torch.empty(nblk, device=p.device, dtype=torch.float64)
    norm = torch.empty(1, device=p.device, dtype=torch.float32)
    _w(2.0 * n, 4.0 * n)
    _call("xm_sumsq_partials_f32", _p(g), n, _p(part), _stream())
    _w(12.0 * n, 28.0 * n)
    _call("xm_clip_adamw_f32", _p(p), _p(g), _p(m), _p(v), n, _p(part), nblk, float(max_norm), float(lr), float(betas[0]),
          float(betas[1]), float(eps), float(weight_decay), int(step), _p(norm), _stream())
    return norm


def clip_adamw_dev_(p, g, m, v, step_dev, lr_dev, weight_decay, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8):
    """clip_adamw_ with the 1-based step count (int64, 1 element) and the learning rate (float32, 1 element) read from
    device memory: the form a CUDA-graph capture of the step needs (kernel arguments are frozen at capture)."""
    _chk(p, g, m, v, lr_dev)
    if step_dev.dtype != torch.int64 or not step_dev.is_cuda:
        raise _lib.XmodalError("step_dev: 1-element int64 CUDA tensor expected")
    n = p.numel()
    nblk = _lib.lib().xm_sumsq_nblk(n)
    part = torch.empty(nblk, device=p.device, dtype=torch.float64)
    norm = torch.empty(1, device=p.device, dtype=torch.float32)
    _w(2.0 * n, 4.0 * n)
    _call("xm_sumsq_partials_f32", _p(g), n, _p(part), _stream())
    _w(12.0 * n, 28.0 * n)
    _call("xm_clip_adamw_dev_f32", _p(p), _p(g), _p(m), _p(v), n, _p(part), nblk, float(max_norm), _p(lr_dev), float(betas[0]),
          float(betas[1]), float(eps), float(weight_decay), _p(step_dev), _p(norm), _stream())
    return norm


# -- seed epoch (include/xmodal_b200.h): device-resident counter folded into every dropout hash ----------------------
def seed_epoch_init() -> None:
    """Allocate the counter on the current device (not capturable; idempotent)."""
    _lib.call("xm_seed_epoch_init")


def seed_epoch_advance() -> None:
    """epoch += 1 on the current stream (capturable: first node of a captured step)."""
    _call("xm_seed_epoch_advance", _stream())


def seed_epoch_set(value: int) -> None:
    seed_epoch_init()
    _call("xm_seed_epoch_set", int(value), _stream())


def seed_epoch_get() -> int:
    seed_epoch_init()
    out = ctypes.c_uint64(0)
    _lib.call("xm_seed_epoch_get", ctypes.byref(out))
    return int(out.value)
