"""Autograd layer over the C-ABI kernels: each class fuses one chain of reference layers
(Conv1d-BatchNorm-GELU-MaxPool-Dropout, Linear-BatchNorm-ReLU-Dropout, Linear-LayerNorm-GELU-Dropout,
similarity + symmetric InfoNCE, ...) into a handful of hand-written kernels, forward and backward.

Conventions: activations between conv blocks are channels-last `(B, T, C)` (see csrc/gemm_engine.cuh);
dropout masks and max-pool argmaxes are recomputed in the backward from a saved 64-bit seed, never
stored; values that feed the next tensor-core contraction are rounded to tf32 when written.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import ops

# ----------------------------------------------------------------------------- context


@dataclass
class ParallelContext:
    """Data-parallel context consulted by the fused blocks.

    group: torch.distributed process group (None = single process).  With `sync_bn` the per-channel
    BatchNorm partial sums are all-reduced so a sharded run reproduces the single-process global
    batch statistics."""

    group: Optional[object] = None
    sync_bn: bool = True
    peer_memory: bool = True  # read the other ranks' embeddings in place over NVLink (symmetric memory)

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if self.active else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if self.active else 0

    @property
    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and (self.group is not None or dist.get_world_size() > 1)


_CTX = ParallelContext()


def parallel_context() -> ParallelContext:
    return _CTX


def set_parallel_context(ctx: ParallelContext) -> None:
    global _CTX
    _CTX = ctx


_seed_counter = 0  # seeds drawn since manual_seed
_base_seed = 0x5EED


def manual_seed(seed: int) -> None:
    """Re-seed the dropout mask stream (counter-based: mask = hash(seed, element index))."""
    global _seed_counter, _base_seed
    _base_seed = int(seed) & 0xFFFFFFFF
    _seed_counter = 0


def next_seed() -> int:
    global _seed_counter
    _seed_counter += 1
    return ((_base_seed << 32) ^ (_seed_counter * 0x9E3779B1) ^ (_CTX.rank << 20)) & 0x7FFFFFFFFFFFFFFF


def seed_state():
    """(base seed, seeds drawn): restore with set_seed_state to re-draw the same seeds."""
    return _base_seed, _seed_counter


def set_seed_state(state) -> None:
    global _seed_counter, _base_seed
    _base_seed, _seed_counter = int(state[0]), int(state[1])


class _PeerReduce:
    """Symmetric-memory workspace of the small fp64 all-reduces (SyncBN statistics): per channel (= issuing CUDA stream:
    the EEG encoder and the fMRI branch run concurrently, each in its own program order) `SLOTS` rotating slots of
    world x ROW doubles plus one sequence flag per (slot, source rank).  ops.peer_allreduce_f64 pushes, publishes, waits
    and sums in one single-CTA kernel -- no communicator, so the two streams never queue behind each other's collectives
    (with one NCCL communicator they did: 32.9 vs 31.1 ms per step on 2 GPUs against a 0.5 ms difference of the
    serialised kernels, profiles/r2_bench_2gpu_v9.json)."""

    CHANNELS, SLOTS, ROW = 4, 4, 1024
    _state = None
    _failed = False

    @classmethod
    def _setup(cls, device):
        import torch.distributed._symmetric_memory as symm
        group = _CTX.group if _CTX.group is not None else dist.group.WORLD
        world = dist.get_world_size(group)
        n_data = cls.CHANNELS * cls.SLOTS * world * cls.ROW
        n_flag = cls.CHANNELS * cls.SLOTS * world
        buf = symm.empty(n_data + n_flag, dtype=torch.float64, device=device)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        torch.cuda.synchronize(device)
        hdl.barrier(channel=0)  # every rank's flags are zero before anyone publishes
        torch.cuda.synchronize(device)
        return {"buf": buf, "hdl": hdl, "ptrs": [int(p) for p in hdl.buffer_ptrs], "world": world, "rank": dist.get_rank(group),
                "n_data": n_data, "seq": {}, "channels": {}}

    @classmethod
    def reduce(cls, t: torch.Tensor):
        """-> the summed tensor, or None when peer memory is unavailable (the caller then uses the collective library)."""
        if cls._failed or not _CTX.peer_memory or not t.is_cuda or t.dtype != torch.float64 or t.numel() > cls.ROW:
            return None
        if cls._state is None:
            try:
                cls._state = cls._setup(t.device)
            except Exception as exc:  # noqa: BLE001 - no P2P mapping on this system
                cls._failed = True
                import warnings
                warnings.warn(f"symmetric memory unavailable ({exc}); SyncBN statistics fall back to NCCL all-reduce")
                return None
        st = cls._state
        sid = torch.cuda.current_stream(t.device).cuda_stream
        ch = st["channels"].setdefault(sid, len(st["channels"]))  # streams are met in the same order on every rank
        if ch >= cls.CHANNELS:
            return None
        seq = st["seq"].get(ch, 0) + 1
        st["seq"][ch] = seq
        world, rank = st["world"], st["rank"]
        slot = (ch * cls.SLOTS + seq % cls.SLOTS) * world  # first (slot, rank) row of this call
        data_off = (slot + rank) * cls.ROW * 8
        flag_off = (st["n_data"] + slot + rank) * 8
        data_dst = [p + data_off for p in st["ptrs"]]
        flag_dst = [p + flag_off for p in st["ptrs"]]
        mine = st["ptrs"][rank]
        out = ops.peer_allreduce_f64(t.reshape(-1), data_dst, flag_dst, mine + slot * cls.ROW * 8,
                                     mine + (st["n_data"] + slot) * 8, cls.ROW, seq)
        return out.view_as(t)


def peer_exchange_counts() -> dict:
    """{channel: calls issued so far} of the SyncBN peer exchange (empty when it is not in use)."""
    st = _PeerReduce._state
    return dict(st["seq"]) if st is not None else {}


def _allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    if _CTX.peer_memory and t.is_cuda:
        out = _PeerReduce.reduce(t)
        if out is not None:
            return out
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_CTX.group)
    return t


def _bn_stats(y, eps, running_mean, running_var, momentum, training, part=None):
    """mean / invstd / count for a (B,T,C) or (B,C) tensor: batch statistics (optionally synchronised
    across ranks) in training, running statistics in eval.  part: the partial sums (rows, C, 2) fp64 of y when its
    producer already accumulated them (conv epilogue)."""
    rows = y.shape[0] * (y.shape[1] if y.dim() == 3 else 1)
    if not training:
        invstd = torch.rsqrt(running_var + eps)
        return running_mean, invstd, float(rows)
    count = float(rows)
    if _CTX.active and _CTX.sync_bn:
        count *= _CTX.world
    if count <= 1:  # torch.nn.functional.batch_norm raises here too (the batch variance is undefined)
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(y.shape)}")
    if part is None:
        part = ops.bn_partial_stats(y)
    if _CTX.active and _CTX.sync_bn:
        part = _allreduce_sum(part.sum(0, keepdim=True))
    mean, invstd = ops.bn_finalize_stats(part, count, eps, running_mean, running_var, momentum)
    return mean, invstd, count


def _bn_backward(dout, y, mean, invstd, gamma, beta, count, act, pool, p, seed, dbp, training, round_out):
    part = ops.bn_act_bwd_reduce(dout, y, mean, invstd, gamma, beta, act, pool, p, seed, dbp)
    if training and _CTX.active and _CTX.sync_bn:
        part_g = _allreduce_sum(part.sum(0, keepdim=True).clone())
        dbeta_g, dgamma_g = ops.bn_bwd_finalize(part_g)  # global sums drive dy
        dbeta, dgamma = ops.bn_bwd_finalize(part)        # local sums are this rank's parameter gradients
    else:
        dbeta, dgamma = ops.bn_bwd_finalize(part)
        dbeta_g, dgamma_g = dbeta, dgamma
    if not training:  # eval-mode BN is an affine map: no statistic terms in dy
        dbeta_g, dgamma_g = torch.zeros_like(dbeta), torch.zeros_like(dgamma)
    dy = ops.bn_act_bwd_apply(dout, y, mean, invstd, gamma, beta, dbeta_g, dgamma_g, count, act, pool, p, seed, dbp,
                              round_out)
    return dy, dgamma, dbeta


# ----------------------------------------------------------------------------- layout
class ToChannelsLast(torch.autograd.Function):
    """(B, C, T) reference layout -> (B, T, C) device layout, rounding to tf32 for the first conv."""

    @staticmethod
    def forward(ctx, x, round_out, split3):
        ctx.C = x.shape[1]
        return ops.to_nwc(x, round_out, split3)

    @staticmethod
    def backward(ctx, g):
        # the same transposing kernel maps a dense (B, T, C) gradient back to (B, C, T); of a channel-stacked split
        # (B, T, 2C) only the first block carries the gradient (ConvBnAct.backward)
        return ops.to_nwc(g[:, :, :ctx.C].contiguous()), None, None


def to_channels_last(x, round_out=True, split3=False):
    """split3: (B, T, 2C) channel-stacked tf32 split [hi | lo], the input of a `precise` conv block (read as [hi | lo | hi])."""
    return ToChannelsLast.apply(x, round_out, split3)


# ----------------------------------------------------------------------------- conv block
class ConvBnAct(torch.autograd.Function):
    """Conv1d("same") -> BatchNorm1d -> act -> [MaxPool1d(2)] -> Dropout on channels-last input.

    enhanced_models_v4.py:128-144, 199-221 ; crossmodal_v4_enhancements.py:822-834, 854-866."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, running_mean, running_var, cfg):
        eps, momentum, act, pool, p, dbp, training, round_out, precise = cfg
        Cout, Cin, taps = w.shape
        x = ops.as_nwc(x)
        xin_cols = x.shape[2]
        wk, wt = ops.conv1d_pack_weight(w)
        if precise:
            # fp32-accurate forward on the tf32 tensor cores: x = xh + xl, w = wh + wl, y = xh wh + xl wh + xh wl as
            # ONE conv over 3 Cin stacked channels [xh | xl | xh] x [wh | wh | wl].  The first two convs of the v4
            # encoder need it: their single-pass operand rounding alone puts 7e-3 .. 1e-2 on five parameter
            # gradients (tools/tf32_floor_by_layer.py, profiles/r2_tf32_floor_by_layer.json); the backward does not.
            if training:  # the conv epilogue also accumulates the BatchNorm batch statistics of y
                y, x, part = ops.conv1d_fwd_precise(x, w, b, stats=True)
            else:
                (y, x), part = ops.conv1d_fwd_precise(x, w, b), None  # x: now the tf32-rounded input, operand of the weight gradient
        elif training:
            y, part = ops.conv1d_fwd(x, wk, b, Cout, stats=True)
        else:
            y, part = ops.conv1d_fwd(x, wk, b, Cout), None
        mean, invstd, count = _bn_stats(y, eps, running_mean, running_var, momentum, training, part)
        seed = next_seed() if (training and p > 0) else 0
        pd = p if training else 0.0
        out = ops.bn_act_fwd(y, mean, invstd, gamma, beta, act, pool, pd, seed, dbp, round_out)
        ctx.save_for_backward(x, y, wt, mean, invstd, gamma, beta)
        ctx.meta = (Cin, taps, count, act, pool, pd, seed, dbp, training, xin_cols)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, wt, mean, invstd, gamma, beta = ctx.saved_tensors
        Cin, taps, count, act, pool, pd, seed, dbp, training, xin_cols = ctx.meta
        dout = ops.as_nwc(dout)  # may be 3 Cout wide (the block fed a precise conv): the kernels read the first Cout columns
        dy, dgamma, dbeta = _bn_backward(dout, y, mean, invstd, gamma, beta, count, act, pool, pd, seed, dbp, training, True)
        dx = None
        if ctx.needs_input_grad[0]:
            if xin_cols == Cin:
                dx = ops.conv1d_dgrad(dy, wt, Cin, round_out=True)
            else:  # the input was a channel-stacked split [hi | lo] (B, T, 2 Cin): its gradient lives in the first block
                dx = ops.empty_pitched((dy.shape[0], dy.shape[1], xin_cols), dy.device)
                ops.conv1d_dgrad(dy, wt, Cin, round_out=True, out=dx[:, :, :Cin])
        # A bias in front of a train-mode BatchNorm has an exactly zero gradient (sum over the batch of the BN
        # input-gradient vanishes identically); the column reduction is only run when BN uses running stats.
        dw, db = ops.conv1d_wgrad(dy, x, taps, need_bias=not training)
        if training:
            db = torch.zeros(dy.shape[2], device=dy.device, dtype=dy.dtype)
        return dx, dw, db, dgamma, dbeta, None, None, None


def _bn_momentum(bn, training) -> float:
    """nn.BatchNorm1d's running-statistics factor for this call: counts the batch, and `momentum=None` means the
    cumulative moving average 1 / num_batches_tracked (torch/nn/modules/batchnorm.py)."""
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    if bn.momentum is not None:
        return float(bn.momentum)
    if training and bn.num_batches_tracked is not None:
        return 1.0 / float(bn.num_batches_tracked.item())
    return 0.0


# fp32-accurate (3-pass) forward of the convolutions the modules mark `precise` (XM_CONV_PRECISE=0: single pass, for A/B)
CONV_PRECISE = os.environ.get("XM_CONV_PRECISE", "1") != "0"


def conv_bn_act(x, conv, bn, act="gelu", pool=0, drop_p=0.0, drop_before_pool=False, training=True, round_out=True,
                precise=False):
    """conv / bn: torch modules used as parameter containers (nn.Conv1d, nn.BatchNorm1d).  `precise`: the conv forward
    runs in the 3-pass mode (see ConvBnAct); its input is either NOT tf32-rounded by its producer or already the
    channel-stacked split [hi | lo] (B, T, 2 Cin).  round_out: False | True (tf32) | 2 (emit that split for a following precise conv)."""
    cfg = (bn.eps, _bn_momentum(bn, training), act, pool, float(drop_p), bool(drop_before_pool),
           bool(training), int(round_out), bool(precise))
    return ConvBnAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg)


# ----------------------------------------------------------------------------- linear blocks
def _tf32(t, already=False):
    """Operand for a tensor-core contraction: rounded to nearest tf32 unless its producer already did.
    (The tensor core would otherwise TRUNCATE the low 13 mantissa bits, a systematic -2^-11 relative bias
    per operand that accumulates across layers; round-to-nearest leaves only zero-mean noise.)"""
    return t if already else ops.round_tf32(t)


class Linear(torch.autograd.Function):
    """y = x @ w^T + b on the tcgen05 GEMM (nn.Linear).

    precise=True  : three tf32 passes over a tripled contraction axis (fp32-accurate; the small,
                    error-sensitive projections: fMRI MLPs, bridge, gates, heads).
    precise=False : one tf32 pass with operands rounded to nearest at their producers (the large
                    transformer projections).  `round_out` rounds y to tf32 (y feeds another contraction
                    directly); `x_rounded` says the producer of x already rounded it."""

    @staticmethod
    def forward(ctx, x, w, b, round_out=False, x_rounded=False, precise=False):
        ctx.has_bias, ctx.precise = b is not None, precise
        if precise:
            ctx.save_for_backward(x, w)
            return ops.linear_fwd_precise(x, w, b)
        xr, wr = _tf32(x, x_rounded), _tf32(w)
        ctx.save_for_backward(xr, wr)
        return ops.linear_fwd(xr, wr, b, round_out=round_out)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.precise:
            dx = ops.linear_dgrad_precise(dy, w) if ctx.needs_input_grad[0] else None
            dw, db = ops.linear_wgrad_precise(dy, x, need_bias=ctx.has_bias)
        else:
            dyr = _tf32(dy)
            dx = ops.linear_dgrad(dyr, w) if ctx.needs_input_grad[0] else None
            dw, db = ops.linear_wgrad(dyr, x, need_bias=ctx.has_bias)
        return dx, dw, db, None, None, None


def linear(x, lin, precise=True):
    return Linear.apply(x, lin.weight, lin.bias, False, False, precise)


class ActDropout(torch.autograd.Function):
    """act -> Dropout (nn.GELU / nn.ReLU / nn.Tanh / nn.Sigmoid followed by nn.Dropout)."""

    @staticmethod
    def forward(ctx, x, act, p, seed):
        ctx.save_for_backward(x)
        ctx.meta = (act, p, seed)
        return ops.act_fwd(x, act, p, seed)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        act, p, seed = ctx.meta
        return ops.act_bwd(g, x, act, p, seed), None, None, None


def act_dropout(x, act, drop_p=0.0, training=True):
    p = float(drop_p) if training else 0.0
    if p == 0.0 and act in (None, "none", 0):
        return x  # identity (a bare nn.Dropout in eval mode / with p = 0)
    return ActDropout.apply(x, act, p, next_seed() if p > 0 else 0)


# Contractions at least this long would run single-pass tf32 in LinearBnAct.  Disabled (every K is "short"):
# measured with tools/acc_probe.py, a single-pass 40 000-d connectivity projection costs 1.2 ms less per step
# but its 3e-4 forward noise is amplified by the BatchNorm-backward cancellation into a 1.4e-2 error of that
# layer's weight gradient (7e-4 in the 3-pass mode).
_PRECISE_MAX_K = 1 << 30
_FUSED_FFN = os.environ.get("XM_FUSED_FFN", "1") != "0"  # A/B switch: 0 = the unfused linear / act / linear chain
_INFONCE_PRECISE_DGRAD = {"0": False, "1": True}.get(os.environ.get("XM_INFONCE_PRECISE_DGRAD", ""), None)
_FUSED_INFONCE_BWD = os.environ.get("XM_FUSED_INFONCE", os.environ.get("XM_FUSED_INFONCE_BWD", "1")) != "0"  # A/B switch: 0 = the unfused lse / grad / split / dgrad GEMM chain


class LinearBnAct(torch.autograd.Function):
    """Linear -> BatchNorm1d -> act -> Dropout on (B, F) rows (fMRI_CODE/fmri_utils.py:26-35, 66-71;
    crossmodal_v4_enhancements.py:696-723, 768-773, 909-914).  The projection runs in the fp32-accurate
    3-pass mode: BatchNorm's backward subtracts batch means (a cancellation that would amplify
    single-pass tf32 noise of these small layers into percent-level gradient errors)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, running_mean, running_var, cfg):
        eps, momentum, act, p, training, prepared = cfg
        # 3-pass (fp32-accurate) projection (see _PRECISE_MAX_K for the single-pass escape hatch)
        precise = prepared or x.shape[1] < _PRECISE_MAX_K
        if precise:
            # one row-stacked tf32 split of the input serves the forward AND the weight gradient (for the 40 000-d
            # connectivity features that is a 2 GB write saved per step); it replaces x among the saved tensors.
            # `prepared`: the producer of x (the connectivity kernel) already wrote that split.
            if not prepared:
                x = ops.linear_precise_prepare(x)
            y = ops.linear_fwd_prepared(x, w, b)
        else:
            x, w = _tf32(x), _tf32(w)
            y = ops.linear_fwd(x, w, b)
        ctx.precise = precise
        mean, invstd, count = _bn_stats(y, eps, running_mean, running_var, momentum, training)
        seed = next_seed() if (training and p > 0) else 0
        pd = p if training else 0.0
        out = ops.bn_act_fwd(y, mean, invstd, gamma, beta, act, 0, pd, seed, False, False)
        ctx.save_for_backward(x, w, y, mean, invstd, gamma, beta)
        ctx.meta = (count, act, pd, seed, training)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, y, mean, invstd, gamma, beta = ctx.saved_tensors
        count, act, pd, seed, training = ctx.meta
        dout = dout.contiguous()
        dy, dgamma, dbeta = _bn_backward(dout, y, mean, invstd, gamma, beta, count, act, 0, pd, seed, False, training,
                                         not ctx.precise)
        if ctx.precise:
            dx = ops.linear_dgrad_precise(dy, w) if ctx.needs_input_grad[0] else None
            dw, db = ops.linear_wgrad_prepared(dy, x, need_bias=not training, K=w.shape[1])
        else:  # x, w were saved tf32-rounded; dy was rounded by the BN backward kernel
            dx = ops.linear_dgrad(dy, w) if ctx.needs_input_grad[0] else None
            dw, db = ops.linear_wgrad(dy, x, need_bias=not training)
        if training:  # bias in front of a train-mode BatchNorm: gradient identically zero
            db = torch.zeros(dy.shape[1], device=dy.device, dtype=dy.dtype)
        return dx, dw, db, dgamma, dbeta, None, None, None


def linear_bn_act(x, lin, bn, act="relu", drop_p=0.0, training=True, prepared=False):
    """prepared: x is the row-stacked tf32 split (3B, K) of the layer input (data, no gradient)."""
    cfg = (bn.eps, _bn_momentum(bn, training), act, float(drop_p), bool(training), bool(prepared))
    return LinearBnAct.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg)


class LinearLnAct(torch.autograd.Function):
    """Linear -> LayerNorm -> act -> Dropout (bridge_utils.py:34-45, 60-66)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, cfg):
        eps, act, p, training = cfg
        y = ops.linear_fwd_precise(x, w, b)
        seed = next_seed() if (training and p > 0) else 0
        pd = p if training else 0.0
        out, mean, rstd = ops.ln_act_fwd(y, gamma, beta, eps, act, pd, seed)
        ctx.save_for_backward(x, w, y, mean, rstd, gamma, beta)
        ctx.meta = (act, pd, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, y, mean, rstd, gamma, beta = ctx.saved_tensors
        act, pd, seed = ctx.meta
        dy, dgamma, dbeta = ops.ln_act_bwd(dout.contiguous(), y, gamma, beta, mean, rstd, act, pd, seed)
        dx = ops.linear_dgrad_precise(dy, w) if ctx.needs_input_grad[0] else None
        dw, db = ops.linear_wgrad_precise(dy, x, need_bias=True)
        return dx, dw, db, dgamma, dbeta, None


def linear_ln_act(x, lin, ln, act="gelu", drop_p=0.0, training=True):
    return LinearLnAct.apply(x, lin.weight, lin.bias, ln.weight, ln.bias, (ln.eps, act, float(drop_p), bool(training)))


# ----------------------------------------------------------------------------- attention core
class SelfAttentionCore(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) -> Dropout -> @ v per (sample, head), on the packed in_proj output
    qkv (B, L, 3d) (nn.MultiheadAttention inside TemporalTransformerBlock, enhanced_models_v4.py:71-73,98).
    Neither scores nor probabilities reach HBM (csrc/attention_fused.cu): only the output and the row
    logsumexp are saved; the backward regenerates the probabilities (and the dropout mask) on chip."""

    @staticmethod
    def forward(ctx, qkv, nhead, p, seed):
        dh = qkv.shape[2] // 3 // nhead
        scale = 1.0 / (dh ** 0.5)
        out, lse = ops.attn_fused_fwd(qkv, nhead, scale, p, seed)
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (nhead, scale, p, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        nhead, scale, p, seed = ctx.meta
        return ops.attn_fused_bwd(_tf32(dout.contiguous()), qkv, out, lse, nhead, scale, p, seed), None, None, None


def self_attention_core(qkv, nhead, drop_p=0.0, training=True):
    p = float(drop_p) if training else 0.0
    return SelfAttentionCore.apply(qkv, nhead, p, next_seed() if p > 0 else 0)


def attention_core_supported(L: int, dh: int) -> bool:
    return ops.attn_supported(L, dh)


class GeneralAttentionCore(torch.autograd.Function):
    """The same attention core for the shapes the fused tcgen05 kernel does not cover (head dim != 32, L > 512) and
    for nn.MultiheadAttention's `attn_mask` (enhanced_models_v4.py:89,98): csrc/attention_general.cu, fp32 SIMT,
    scores recomputed in the backward from the saved logsumexp.  `mask`: additive fp32 (L, L) | (B*H, L, L) | None."""

    @staticmethod
    def forward(ctx, qkv, mask, nhead, p, seed):
        scale = 1.0 / ((qkv.shape[2] // 3 // nhead) ** 0.5)
        out, lse = ops.attn_general_fwd(qkv, nhead, scale, mask, p, seed)
        ctx.save_for_backward(qkv, lse, mask) if mask is not None else ctx.save_for_backward(qkv, lse)
        ctx.meta = (nhead, scale, p, seed, mask is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        nhead, scale, p, seed, has_mask = ctx.meta
        qkv, lse = ctx.saved_tensors[:2]
        mask = ctx.saved_tensors[2] if has_mask else None
        return ops.attn_general_bwd(dout.contiguous(), qkv, lse, nhead, scale, mask, p, seed), None, None, None, None


def general_attention_core(qkv, nhead, mask=None, drop_p=0.0, training=True):
    p = float(drop_p) if training else 0.0
    return GeneralAttentionCore.apply(qkv, mask, nhead, p, next_seed() if p > 0 else 0)


def general_attention_supported(L: int, dh: int) -> bool:
    return ops.attn_general_supported(L, dh)


def additive_attention_mask(mask, B: int, nhead: int, L: int):
    """nn.MultiheadAttention's `attn_mask` -> the additive fp32 mask of the attention kernels: a boolean (or uint8)
    mask marks positions that may NOT be attended with True (-> -inf); a float mask is added to the scores as is.
    Accepted shapes, as in torch: (L, L) or (B * num_heads, L, L)."""
    if mask is None:
        return None
    if tuple(mask.shape) not in ((L, L), (B * nhead, L, L)):
        raise ValueError(f"attn_mask of shape {tuple(mask.shape)}: expected ({L}, {L}) or ({B * nhead}, {L}, {L})")
    if mask.dtype in (torch.bool, torch.uint8):
        return torch.zeros(mask.shape, device=mask.device, dtype=torch.float32).masked_fill_(mask.bool(), float("-inf"))
    if not mask.is_floating_point():
        raise ValueError(f"attn_mask dtype {mask.dtype}: only bool, uint8 and floating-point masks are defined")
    return mask.to(torch.float32)


class LayerNormAct(torch.autograd.Function):
    """LayerNorm -> act -> Dropout over (M, D) rows for any D (nn.LayerNorm; the fused residual kernels cover only
    D % 128 == 0)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, cfg):
        eps, act, p, seed = cfg
        out, mean, rstd = ops.ln_act_fwd(x, gamma, beta, eps, act, p, seed)
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.meta = (act, p, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        act, p, seed = ctx.meta
        dx, dgamma, dbeta = ops.ln_act_bwd(dout.contiguous(), x, gamma, beta, mean, rstd, act, p, seed)
        return dx, dgamma, dbeta, None


def layer_norm(x, ln, act=None, drop_p=0.0, training=True):
    """ln: nn.LayerNorm used as a parameter container; x: (..., D)."""
    p = float(drop_p) if training else 0.0
    shape = x.shape
    out = LayerNormAct.apply(x.reshape(-1, shape[-1]), ln.weight, ln.bias, (ln.eps, act, p, next_seed() if p > 0 else 0))
    return out.view(shape)


# ----------------------------------------------------------------------------- transformer tail
class TransformerTail(torch.autograd.Function):
    """PositionalEncoding -> N x TemporalTransformerBlock -> mean over time, as ONE function over the token
    matrix (EEG_CODE/enhanced_models_v4.py:44-55, 89-107, 161-163), forward and hand-written backward.

    Per block: [residual add + Dropout + LayerNorm] fused kernel -> in_proj GEMM -> attention core ->
    out_proj GEMM -> [residual add + Dropout + LayerNorm] -> linear1 GEMM -> GELU+Dropout -> linear2 GEMM;
    the add of a block's FFN output is fused into the NEXT block's LayerNorm kernel (or into the final mean).
    Every GEMM operand is rounded to tf32 by the kernel that produces it; dropout masks are regenerated from
    seeds in the backward.  params: 12 tensors per block in the order
    (norm1.w, norm1.b, in_proj_weight, in_proj_bias, out_proj.w, out_proj.b, norm2.w, norm2.b, linear1.w,
    linear1.b, linear2.w, linear2.b)."""

    NP = 12

    @staticmethod
    def forward(ctx, h0, pe, cfg, *params):
        nhead, p, eps, act, training = cfg
        p = float(p) if training else 0.0
        B, L, D = h0.shape
        M = B * L
        nl = len(params) // TransformerTail.NP
        scale = 1.0 / ((D // nhead) ** 0.5)
        seed = (lambda: next_seed()) if p > 0 else (lambda: 0)
        fused_ffn = _FUSED_FFN and nl > 0 and ops.ffn_fused_supported(D, params[8].shape[0], act)
        x = h0.reshape(M, D)
        saved, meta = [], []
        pend, pend_seed = None, 0
        seed_pe = seed()
        for l in range(nl):
            n1w, n1b, wqkv, bqkv, wo, bo, n2w, n2b, w1, b1, w2, b2 = params[l * 12:(l + 1) * 12]
            if l == 0:
                x1, h1, m1, r1 = ops.resid_ln_fwd(x, None, n1w, n1b, eps, p, seed_pe, pe=pe, L=L)
            else:
                x1, h1, m1, r1 = ops.resid_ln_fwd(x, pend, n1w, n1b, eps, p, pend_seed)
            wqkv_r, wo_r, w1_r, w2_r = (ops.round_tf32(w) for w in (wqkv, wo, w1, w2))
            qkv = ops.linear_fwd(h1, wqkv_r, bqkv, round_out=True)
            s_attn = seed()
            att, lse = ops.attn_fused_fwd(qkv.view(B, L, 3 * D), nhead, scale, p, s_attn, round_out=True)
            ao = ops.linear_fwd(att.view(M, D), wo_r, bo)
            s_ao = seed()
            x2, h2, m2, r2 = ops.resid_ln_fwd(x1, ao, n2w, n2b, eps, p, s_ao)
            del ao
            s_g = seed()
            if fused_ffn:  # the (M, hidden) activations stay on chip; the backward recomputes them from h2
                f2 = ops.ffn_fused_fwd(h2, w1_r, b1, w2_r, b2, act, p, s_g)
                f1, g = b1, None
            else:
                f1 = ops.linear_fwd(h2, w1_r, b1)
                g = ops.act_fwd(f1, act, p, s_g, round_out=True)
                f2 = ops.linear_fwd(g, w2_r, b2)
            saved += [x1, h1, m1, r1, qkv, lse, att, x2, h2, m2, r2, f1, g, wqkv_r, wo_r, w1_r, w2_r, n1w, n2w]
            meta.append((s_attn, s_ao, s_g, pend_seed if l > 0 else seed_pe))
            x, pend, pend_seed = x2, f2, seed()
        out = ops.resid_seqmean_fwd(x.view(B, L, D), pend.view(B, L, D), p, pend_seed)
        ctx.save_for_backward(*saved)
        ctx.meta = (B, L, D, nl, nhead, p, act, scale, meta, pend_seed, fused_ffn)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, L, D, nl, nhead, p, act, scale, meta, last_seed, fused_ffn = ctx.meta
        M = B * L
        sv = ctx.saved_tensors
        dxs, df2 = ops.resid_seqmean_bwd(dout.contiguous(), L, p, last_seed)
        dxs, df2 = dxs.view(M, D), df2.view(M, D)
        db2 = ops.colsum(df2)  # bias gradient of the last block's linear2
        grads = [None] * (nl * 12)
        dh0 = None
        for l in reversed(range(nl)):
            (x1, h1, m1, r1, qkv, lse, att, x2, h2, m2, r2, f1, g, wqkv_r, wo_r, w1_r, w2_r, n1w, n2w) = sv[l * 19:(l + 1) * 19]
            s_attn, s_ao, s_g, s_in = meta[l]
            if fused_ffn:
                # one pass recomputes the hidden activations and emits both weight-gradient operands, dh2 and db1
                g, df1, dh2, db1 = ops.ffn_fused_dgrad(h2, df2, w1_r, f1, w2_r.t().contiguous(), w1_r.t().contiguous(), act, p, s_g)
                dw2, _ = ops.linear_wgrad(df2, g, need_bias=False)
                dw1, _ = ops.linear_wgrad(df1, h2, need_bias=False)
                del g, df1
            else:
                dg = ops.linear_dgrad(df2, w2_r)
                dw2, _ = ops.linear_wgrad(df2, g, need_bias=False)
                df1, db1 = ops.act_bwd_colsum(dg, f1, act, p, s_g, round_out=True)  # bias gradient in the same pass
                del dg
                dh2 = ops.linear_dgrad(df1, w1_r)
                dw1, _ = ops.linear_wgrad(df1, h2, need_bias=False)
                del df1
            # bias gradients of out_proj / the previous linear2 come out of the fused LayerNorm-backward pass
            dx1, dao, dn2w, dn2b, dbo = ops.resid_ln_bwd(dh2, dxs, x2, n2w, m2, r2, p, s_ao)
            datt = ops.linear_dgrad(dao, wo_r, round_out=True)
            dwo, _ = ops.linear_wgrad(dao, att.view(M, D), need_bias=False)
            # the in-projection's bias gradient (column sums of dqkv) is accumulated where dq / dk / dv leave tensor memory
            dqkv, dbqkv = ops.attn_fused_bwd(datt.view(B, L, D), qkv.view(B, L, 3 * D), att.view(B, L, D), lse, nhead, scale, p,
                                             s_attn, round_out=True, need_bias=True)
            dqkv = dqkv.view(M, 3 * D)
            dh1 = ops.linear_dgrad(dqkv, wqkv_r)
            dwqkv, _ = ops.linear_wgrad(dqkv, h1, need_bias=False)
            need_in = l > 0 or ctx.needs_input_grad[0]
            dxs, dprev, dn1w, dn1b, dbprev = ops.resid_ln_bwd(dh1, dx1, x1, n1w, m1, r1, p, s_in, need_da=need_in)
            grads[l * 12:(l + 1) * 12] = [dn1w, dn1b, dwqkv, dbqkv, dwo, dbo, dn2w, dn2b, dw1, db1, dw2, db2]
            if l > 0:
                df2, db2 = dprev, dbprev  # gradient (and bias gradient) of the previous block's FFN branch
            else:
                dh0 = None if dprev is None else dprev.view(B, L, D)  # d/d(conv output): Dropout(x + pe) backward
        return (dh0, None, None, *grads)


def transformer_tail_supported(L: int, D: int, nhead: int, act: str) -> bool:
    return (D % nhead == 0 and ops.attn_supported(L, D // nhead) and ops.resid_ln_supported(D)
            and act in ("gelu", "relu"))


# ----------------------------------------------------------------------------- pooling
class SeqMean(torch.autograd.Function):
    """AdaptiveAvgPool1d(1) + Flatten on channels-last input: (B, T, C) -> (B, C)."""

    @staticmethod
    def forward(ctx, x):
        ctx.T = x.shape[1]
        return ops.seqmean(ops.as_nwc(x))

    @staticmethod
    def backward(ctx, g):
        return ops.seqmean_bwd(g, ctx.T)


def seq_mean(x):
    return SeqMean.apply(x)


# ----------------------------------------------------------------------------- fMRI ROI aggregation
def roi_meanstd(x):
    """fMRI_CODE/fmri_utils.py:140-147 on the device; inputs are data (no gradient)."""
    return ops.roi_meanstd(x)


# ----------------------------------------------------------------------------- similarity + InfoNCE
class _AllGatherRows(object):
    @staticmethod
    def gather(t: torch.Tensor) -> torch.Tensor:
        if not _CTX.active:
            return t
        out = torch.empty(_CTX.world * t.shape[0], *t.shape[1:], device=t.device, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t.contiguous(), group=_CTX.group)
        return out


class _PeerShards:
    """Symmetric-memory buffers for the tf32-split unit embeddings: every rank writes its (Bl, 3D) shard into
    its own buffer, which all ranks of the node can read through NVLink; the InfoNCE kernels then take the
    list of per-rank base pointers instead of an all-gathered copy (fused all-gather + GEMM).  Two buffer
    sets alternate so a forward may run before the previous backward has been consumed."""

    _cache = {}
    _failed = False

    @classmethod
    def get(cls, rows: int, cols: int, device):
        if cls._failed or not (_CTX.active and _CTX.peer_memory) or device.type != "cuda" or rows % 128 != 0:
            return None
        key = (rows, cols, device.index)
        ent = cls._cache.get(key)
        if ent is None:
            try:
                import torch.distributed._symmetric_memory as symm
                group = _CTX.group if _CTX.group is not None else dist.group.WORLD
                sets = []
                for _ in range(2):
                    bufs = [symm.empty(rows, cols, dtype=torch.float32, device=device) for _ in range(2)]
                    hdls = [symm.rendezvous(b, group) for b in bufs]
                    sets.append((bufs, hdls))
                ent = {"sets": sets, "turn": 0}
                cls._cache[key] = ent
            except Exception as exc:  # noqa: BLE001 - no P2P mapping on this system: keep the NCCL all-gather path
                cls._failed = True
                import warnings
                warnings.warn(f"symmetric memory unavailable ({exc}); InfoNCE falls back to NCCL all-gather")
                return None
        ent["turn"] ^= 1
        return ent["sets"][ent["turn"]]


class Cross2Attention(torch.autograd.Function):
    """The bridge head's cross-attention core: one query per sample over the two-token sequence [eeg, fmri]
    (bridge_utils.py:74-83), forward and backward in one launch each; returns (out, dropped weights (B, H, 2))."""

    @staticmethod
    def forward(ctx, q, kv, nhead, p, training):
        seed = next_seed() if (training and p > 0) else 0
        pd = p if training else 0.0
        out, att = ops.cross2_attn_fwd(q, kv, nhead, pd, seed)
        ctx.save_for_backward(q, kv)
        ctx.meta = (nhead, pd, seed)
        ctx.mark_non_differentiable(att)
        return out, att

    @staticmethod
    def backward(ctx, dout, _datt):
        q, kv = ctx.saved_tensors
        nhead, pd, seed = ctx.meta
        dq, dkv = ops.cross2_attn_bwd(dout.contiguous(), q, kv, nhead, pd, seed)
        return dq, dkv, None, None, None


def cross2_attention(q, kv, nhead, p=0.0, training=False):
    return Cross2Attention.apply(q.contiguous(), kv.contiguous(), nhead, p, training)


def _infonce_lse_pair(e3, f3, e3_all, f3_all, inv_tau, off):
    """(lse of my e x all f, lse of my f x all e, positives): one fused launch where the shapes allow, else two GEMMs."""
    if _FUSED_INFONCE_BWD and ops.infonce_bwd_fused_supported(e3.shape[0], e3_all.shape[0], e3.shape[1] // 3, off):
        return ops.infonce_lse_fused(e3, f3, e3_all, f3_all, inv_tau, off)
    lse_ef, diag = ops.infonce_lse(e3, f3_all, inv_tau, off)
    lse_fe, _ = ops.infonce_lse(f3, e3_all, inv_tau, off)
    return lse_ef, lse_fe, diag


class SymmetricInfoNCE(torch.autograd.Function):
    """L = 1/(2B) sum_i [lse_j S_ij - S_ii] + [lse_j S_ji - S_ii],  S = norm(e) norm(f)^T / tau, over the
    GLOBAL batch: each rank holds B/G rows of e and f; the similarity kernels read the other ranks' unit
    embeddings tile by tile straight from peer HBM over NVLink (or from an NCCL all-gathered copy when peer
    mapping is unavailable), the two logsumexp vectors are all-gathered, and every rank produces the exact
    gradient of the global loss for its own rows -- no reduce-scatter in the backward.
    S is never materialised in the forward; the backward materialises the (B/G x B) softmax-gradient blocks
    that feed the two dgrad GEMMs.  Returns this rank's share of the loss (sum over ranks = global loss)."""

    @staticmethod
    def forward(ctx, e, f, temperature):
        inv_tau = 1.0 / float(temperature)
        Bl, D = e.shape
        peers = _PeerShards.get(Bl, 3 * D, e.device)
        # fp32 unit vectors + their 3-way tf32 splits: the similarity contraction is fp32-accurate
        if peers is not None:
            (e3, f3), (he, hf) = peers
            en, _, einv = ops.l2norm_split_fwd(e, 0, xs_out=e3)
            fn, _, finv = ops.l2norm_split_fwd(f, 1, xs_out=f3)
            hf.barrier(channel=0)  # every rank's shards are written before anyone reads them
            e_ptrs, f_ptrs = list(he.buffer_ptrs), list(hf.buffer_ptrs)
            world, off = _CTX.world, _CTX.rank * Bl
            Bg = world * Bl
            if Bl <= 4 * 128:
                # few row tiles: read the remote tiles inside the GEMMs (fused all-gather + contraction)
                lse_ef, diag = ops.infonce_lse_peers(e3, f_ptrs, Bl, inv_tau, off)
                lse_fe, _ = ops.infonce_lse_peers(f3, e_ptrs, Bl, inv_tau, off)
                ctx.peers = (e_ptrs, f_ptrs, Bl)
                ctx.save_for_backward(en, fn, einv, finv, e3, f3, lse_ef, lse_fe)
            else:
                # many row tiles: each would re-fetch every remote tile over NVLink (remote memory is not cached
                # in the local L2), so gather once through the peer mappings and contract against the local copy
                e3_all = ops.peer_gather(e_ptrs, Bl, 3 * D, e.device)
                f3_all = ops.peer_gather(f_ptrs, Bl, 3 * D, e.device)
                lse_ef, lse_fe, diag = _infonce_lse_pair(e3, f3, e3_all, f3_all, inv_tau, off)
                ctx.peers = None
                ctx.save_for_backward(en, fn, einv, finv, e3, f3, lse_ef, lse_fe, e3_all, f3_all)
        else:
            en, e3, einv = ops.l2norm_split_fwd(e, 0)
            fn, f3, finv = ops.l2norm_split_fwd(f, 1)
            e3_all = _AllGatherRows.gather(e3)
            f3_all = _AllGatherRows.gather(f3)
            Bg = e3_all.shape[0]
            off = _CTX.rank * Bl if _CTX.active else 0
            lse_ef, lse_fe, diag = _infonce_lse_pair(e3, f3, e3_all, f3_all, inv_tau, off)
            ctx.peers = None
            ctx.save_for_backward(en, fn, einv, finv, e3, f3, lse_ef, lse_fe, e3_all, f3_all)
        loss = (0.5 / Bg) * ((lse_ef - diag).sum() + (lse_fe - diag).sum())
        ctx.meta = (inv_tau, off, Bg)
        return loss

    @staticmethod
    def backward(ctx, g):
        inv_tau, off, Bg = ctx.meta
        sv = ctx.saved_tensors
        en, fn, einv, finv, e3, f3, lse_ef, lse_fe = sv[:8]
        D = en.shape[1]
        lse_ef_all = _AllGatherRows.gather(lse_ef)
        lse_fe_all = _AllGatherRows.gather(lse_fe)
        coef = 0.5 * inv_tau / Bg
        # dE = G f_n, dF = G' e_n: either one tf32 pass with G rounded in its producer's epilogue, or the fp32-accurate
        # 3-pass product (G kept in fp32 and split on the fly).  Measured at batch 2048 (tools/full_scale_parity.py)
        # the 3-pass variant lowers the median parameter-gradient error from 5.4e-4 to 3.0e-4 (fMRI-side tensors from
        # ~7e-4 to ~1e-4; the maximum, 9.3e-3 on the first conv layer, is the tf32 floor either way), but its split of
        # the (local batch x GLOBAL batch) matrix costs ~2 ms per step at 8 x 4096.  Policy: precise while the global
        # batch is <= 8192 (0.25 ms at 4096), single pass beyond; XM_INFONCE_PRECISE_DGRAD=0/1 forces either.
        Ng_all = lse_fe_all.shape[0]
        prec = (Ng_all <= 8192) if _INFONCE_PRECISE_DGRAD is None else _INFONCE_PRECISE_DGRAD
        if ctx.peers is not None:
            e_ptrs, f_ptrs, Bl = ctx.peers
            G1 = ops.infonce_grad_peers(e3, f_ptrs, Bl, lse_ef, lse_fe_all, inv_tau, off, coef, not prec)  # my e x all f
            G2 = ops.infonce_grad_peers(f3, e_ptrs, Bl, lse_fe, lse_ef_all, inv_tau, off, coef, not prec)  # my f x all e
            if prec:  # small here (Bl <= 512): local copies for the split product
                f3_all = ops.peer_gather(f_ptrs, Bl, 3 * D, en.device)
                e3_all = ops.peer_gather(e_ptrs, Bl, 3 * D, en.device)
            else:
                den = ops.linear_dgrad_peers(G1, f_ptrs, Bl, D, 3 * D)
                dfn = ops.linear_dgrad_peers(G2, e_ptrs, Bl, D, 3 * D)
        else:
            e3_all, f3_all = sv[8], sv[9]
            if _FUSED_INFONCE_BWD and ops.infonce_bwd_fused_supported(e3.shape[0], Ng_all, D, off):
                # both (local x global) gradient blocks are formed and contracted in tensor memory, never written
                den, dfn = ops.infonce_bwd_fused(e3, f3, e3_all, f3_all, lse_ef, lse_fe, lse_ef_all, lse_fe_all, inv_tau,
                                                 off, coef, prec)
                return ops.l2norm_bwd(den, en, einv) * g, ops.l2norm_bwd(dfn, fn, finv) * g, None
            G1 = ops.infonce_grad(e3, f3_all, lse_ef, lse_fe_all, inv_tau, off, coef, not prec)
            G2 = ops.infonce_grad(f3, e3_all, lse_fe, lse_ef_all, inv_tau, off, coef, not prec)
            if not prec:  # the first D columns of a split are the tf32-rounded unit vectors
                den = ops.linear_dgrad(G1, f3_all[:, :D])
                dfn = ops.linear_dgrad(G2, e3_all[:, :D])
        if prec:
            den = ops.infonce_dgrad(G1, f3_all, 1)  # f3 = [hi | hi | lo]
            dfn = ops.infonce_dgrad(G2, e3_all, 0)  # e3 = [hi | lo | hi]
        de = ops.l2norm_bwd(den, en, einv) * g
        df = ops.l2norm_bwd(dfn, fn, finv) * g
        return de, df, None


def symmetric_infonce(e, f, temperature=0.07):
    return SymmetricInfoNCE.apply(e.contiguous(), f.contiguous(), temperature)


def similarity_matrix(e, f, temperature=0.07):
    """Materialised S = normalize(e) @ normalize(f)^T / temperature (no gradient; inspection / retrieval)."""
    _, e3, _ = ops.l2norm_split_fwd(e.detach().contiguous(), 0)
    _, f3, _ = ops.l2norm_split_fwd(f.detach().contiguous(), 1)
    return ops.similarity(e3, f3, 1.0 / float(temperature))
