"""The paired EEG/fMRI cross-modal training step and its data-parallel driver.

    raw EEG recordings (R, C, n) --window gather--> (B, T, C) --EnhancedERPEncoder--> (B, 128) --eeg_proj--+
    ROI series (B, TR, ROI) --mean/std--> (B, 2*ROI) --+                                                   |--> symmetric
    connectivity (B, ROI*ROI) -------------------------+--> fMRIFusionNet features (B, 64) --fmri_proj----+    InfoNCE

followed by the reference's step recipe (zero_grad -> backward -> clip_grad_norm_(1.0) -> AdamW,
_test_bridge.py:775-788, lr / weight decay of :63-64,869).

Data parallel: one process per GPU, batch sharded; the InfoNCE function all-gathers the normalised
embeddings (global negatives) and returns exact local gradients of the global loss; BatchNorm
partial sums are all-reduced (SyncBN) so a sharded run equals the single-process global-batch run;
parameter gradients live in ONE flat bucket that is all-reduced (SUM) once per step.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import fmri_utils, ops
from . import functional as XF
from .modules import EEGfMRIBridgeFusionNet, EnhancedERPEncoder, LiteERPEncoder, fMRIFusionNet


class PairedBridgeModel(nn.Module):
    """EEG encoder + fMRI net + bridge projections trained with the symmetric InfoNCE loss."""

    def __init__(self, eeg_channels: int = 64, n_roi: int = 200, conn_dim: Optional[int] = None, eeg_hidden: int = 128,
                 fmri_hidden: int = 64, bridge_dim: int = 128, dropout: float = 0.3, fmri_dropout: float = 0.4,
                 encoder: str = "v4", temperature: float = 0.07):
        super().__init__()
        conn_dim = n_roi * n_roi if conn_dim is None else conn_dim
        if encoder == "v4":
            self.eeg_encoder = EnhancedERPEncoder(eeg_channels, eeg_hidden, 2, 4, dropout)
        elif encoder == "lite":
            self.eeg_encoder = LiteERPEncoder(eeg_channels, eeg_hidden, dropout)
        else:
            raise ValueError(f"unknown encoder {encoder!r}")
        self.fmri_net = fMRIFusionNet(2 * n_roi, conn_dim, fmri_hidden, 2, fmri_dropout)
        self.bridge = EEGfMRIBridgeFusionNet(eeg_hidden, fmri_hidden, bridge_dim, 2, 4, dropout)
        self.temperature = temperature

    def contrastive_parameters(self) -> List[nn.Parameter]:
        """Parameters the InfoNCE step reaches (the supervised heads get no gradient, exactly as
        parameters with grad=None are skipped by the reference's optimizer)."""
        mods: Iterable[nn.Module] = (self.eeg_encoder, self.fmri_net.activation_encoder,
                                     self.fmri_net.connectivity_encoder, self.fmri_net.fusion,
                                     self.bridge.eeg_proj, self.bridge.fmri_proj)
        ps = [p for m in mods for p in m.parameters()]
        return ps + [self.fmri_net.activation_weight, self.fmri_net.connectivity_weight]

    # The fMRI branch (ROI aggregation, the connectivity kernel, two MLP encoders in the 3-pass mode, fusion: ~150 mostly
    # small launches) does not depend on the EEG encoder: on a side stream it fills the tails of the encoder's big
    # kernels, in the forward and -- autograd replays every node on the stream of its forward -- in the backward.
    # Measured on B200 (profiles/r2_bench_overlap.txt): 36.27 -> 35.20 ms per step on 1 GPU; SyncBN collectives issued
    # from both streams keep their order (tests/dp_gpu_check.py).  XM_OVERLAP_BRANCHES=0 serialises the branches
    # (bench.py does so for its per-kernel CUDA-event pass, whose rates must be those of kernels running alone).
    overlap_branches = os.environ.get("XM_OVERLAP_BRANCHES", "1") == "1"

    def _fmri_features(self, roi_series, conn):
        act = fmri_utils.aggregate_roi_timeseries(roi_series, "both")
        if conn is None:  # functional connectivity derived on the device from the same ROI series
            # (the kernel can also write the row-stacked tf32 split the 3-pass projection consumes, `prepared=True`;
            # measured on B200 that costs more than the separate split pass: 1.50 vs 0.79 + 0.51 ms, its mirrored
            # stores then scatter over three planes -- tools/corr_bench.py)
            conn = fmri_utils.connectivity_from_timeseries(roi_series)
        return self.fmri_net.features(act, conn)

    def embed(self, eeg: torch.Tensor, roi_series: torch.Tensor, conn: Optional[torch.Tensor] = None,
              eeg_channels_last: bool = False):
        """conn: (B, conn_dim) connectivity features, or None to derive them as the flattened ROI x ROI correlation
        matrix of `roi_series` on the device (conn_dim must then be n_roi ** 2)."""
        if not (self.overlap_branches and eeg.is_cuda):
            eeg_feat = self.eeg_encoder(eeg, channels_last=eeg_channels_last)
            return self.bridge.project(eeg_feat, self._fmri_features(roi_series, conn))
        cur = torch.cuda.current_stream(eeg.device)
        side = getattr(self, "_side_stream", None)
        if side is None or side.device != eeg.device:
            side = self._side_stream = torch.cuda.Stream(eeg.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fmri_feat = self._fmri_features(roi_series, conn)
        eeg_feat = self.eeg_encoder(eeg, channels_last=eeg_channels_last)
        cur.wait_stream(side)
        for t in (roi_series, conn):
            if t is not None:
                t.record_stream(side)
        fmri_feat.record_stream(cur)
        return self.bridge.project(eeg_feat, fmri_feat)

    def forward(self, eeg, roi_series, conn=None, eeg_channels_last: bool = False) -> torch.Tensor:
        e, f = self.embed(eeg, roi_series, conn, eeg_channels_last)
        return XF.symmetric_infonce(e, f, self.temperature)


def init_distributed(backend: str = "nccl") -> XF.ParallelContext:
    """One process per GPU (torchrun env).  Returns the active context (a no-op context when WORLD_SIZE=1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    ctx = XF.ParallelContext(group=None, sync_bn=True)
    XF.set_parallel_context(ctx)
    return ctx


class PairedTrainer:
    """zero_grad -> forward -> backward -> [all-reduce] -> clip_grad_norm_ -> AdamW (_test_bridge.py:775-788, :869),
    on flat parameter / gradient / moment buffers."""

    def __init__(self, model: PairedBridgeModel, lr: float = 1e-4, weight_decay: float = 1e-4, grad_clip: float = 1.0,
                 window: Optional[int] = None, hop: Optional[int] = None):
        self.model = model
        self.grad_clip = grad_clip
        self.window, self.hop = window, hop
        params = model.contrastive_parameters()
        # flat layout [early | late]: the fMRI branch finishes its backward long before the EEG encoder does (it runs on
        # a side stream), so its gradients -- 10.3 of the 11.6 M parameters -- are all-reduced while the encoder's
        # backward is still running; only the small late bucket stays on the critical path.
        early = {id(p) for p in model.fmri_net.parameters()}
        self.params = [p for p in params if id(p) in early] + [p for p in params if id(p) not in early]
        self._early_params = [p for p in self.params if id(p) in early]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        # Parameters, gradients and both AdamW moments live in ONE flat fp32 buffer each (every tensor 16-B aligned in
        # it): one gradient collective, and norm + clip + update are two launches (ops.clip_adamw_) instead of the
        # foreach kernels of clip_grad_norm_ / torch.optim.AdamW.  The module's parameters become views of the flat
        # buffer, so state_dict() / load_state_dict() keep working on the same storage.
        offs, o = [], 0
        for p in self.params:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self._n_early = offs[len(self._early_params)] if len(self._early_params) < len(self.params) else o
        self.flat_param = torch.zeros(o, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(o, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(o, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(o, device=dev, dtype=torch.float32)
        self._offs = offs
        self._grad_views = []
        for p, o in zip(self.params, offs):
            view = self.flat_param[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            # Gradients: autograd hands every parameter its own tensor (p.grad is None during backward, so nothing is
            # accumulated); ONE gather launch per bucket copies them into the flat buffer (ops.gather_flat_) and p.grad
            # then points at its range of it.  Presetting p.grad to the views instead costs one tiny add kernel per
            # parameter and step (95 of ~300 launches, profiles/r2_launches_v9.md).
            self._grad_views.append(self.flat_grad[o:o + p.numel()].view_as(p))
            p.grad = None
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, (0.9, 0.999), 1e-8
        self.step_count = 0
        self._step_dev = self._lr_dev = None  # set by use_device_state()
        self._lr_on_device = None
        self.last_grad_norm = None  # pre-clip total norm of the last step (device scalar)
        self.ctx = XF.parallel_context()
        self._stage = {}
        self._early_left = 0
        self._early_event = None
        self._comm_stream = None
        if self.ctx.active and dev.type == "cuda" and self._early_params:
            self._early_event = torch.cuda.Event()
            self._comm_stream = torch.cuda.Stream(dev)
            for p in self._early_params:
                p.register_post_accumulate_grad_hook(self._early_grad_ready)

    def _gather(self, lo: int, hi: int) -> None:
        """Parameters [lo, hi): their freshly produced gradient tensors -> the flat buffer (one launch), on the current
        stream; p.grad then aliases the buffer."""
        ps = self.params[lo:hi]
        have = [(p.grad, o) for p, o in zip(ps, self._offs[lo:hi]) if p.grad is not None]
        for p, v in zip(ps, self._grad_views[lo:hi]):
            if p.grad is None:
                v.zero_()  # a parameter the loss did not reach
        if have:
            ops.gather_flat_(self.flat_grad, [g for g, _ in have], [o for _, o in have])
        for p, v in zip(ps, self._grad_views[lo:hi]):
            p.grad = v

    def _early_grad_ready(self, _param) -> None:
        """Runs on the stream that produced the gradient: after the last early parameter, gather the early bucket there
        and mark it ready."""
        self._early_left -= 1
        if self._early_left == 0:
            self._gather(0, len(self._early_params))
            self._early_event.record(torch.cuda.current_stream())

    def _allreduce_gradients(self) -> None:
        g = self.ctx.group
        if self._early_event is None or self._early_left != 0:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=g)
            return
        self._comm_stream.wait_event(self._early_event)
        with torch.cuda.stream(self._comm_stream):
            work = dist.all_reduce(self.flat_grad[:self._n_early], op=dist.ReduceOp.SUM, group=g, async_op=True)
        dist.all_reduce(self.flat_grad[self._n_early:], op=dist.ReduceOp.SUM, group=g)
        work.wait()  # the current stream continues after the early bucket has arrived

    # -- device-resident step ---------------------------------------------------------------
    def step(self, eeg: torch.Tensor, roi_series: torch.Tensor, conn: Optional[torch.Tensor] = None) -> torch.Tensor:
        """eeg: raw recordings (R, C, n) when a window spec was given (gathered on the device into
        B = R*n_win channels-last windows), else windows (B, C, T); conn None: connectivity derived on the device
        from roi_series.  Returns the loss (device scalar, this rank's share of the global loss)."""
        for p in self.params:
            p.grad = None
        if self.window is not None:
            split = bool(getattr(self.model.eeg_encoder, "wants_unrounded_input", False))  # 3-pass first conv
            x = ops.window_gather(eeg, self.window, self.hop or self.window, channels_last=True, round_out=not split,
                                  split3=split)
            loss = self.model(x, roi_series, conn, eeg_channels_last=True)
        else:
            loss = self.model(eeg, roi_series, conn)
        self._early_left = len(self._early_params)
        loss.backward()
        early_done = self._early_event is not None and self._early_left == 0
        self._gather(len(self._early_params) if early_done else 0, len(self.params))
        if self.ctx.active:
            self._allreduce_gradients()
        self.step_count += 1
        if self._step_dev is None:
            self.last_grad_norm = ops.clip_adamw_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq,
                                                  self.step_count, self.lr, self.weight_decay, self.grad_clip or 0.0,
                                                  self.betas, self.eps)
        else:  # device-resident step count / learning rate (a captured step replays with the values left there)
            if not torch.cuda.is_current_stream_capturing():
                self._sync_lr()
            self._step_dev.add_(1)
            self.last_grad_norm = ops.clip_adamw_dev_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq,
                                                      self._step_dev, self._lr_dev, self.weight_decay,
                                                      self.grad_clip or 0.0, self.betas, self.eps)
        return loss.detach()

    # -- device-resident optimizer scalars + CUDA-graph capture ----------------------------------------
    def use_device_state(self) -> None:
        """Move the AdamW step count and the learning rate into device memory (idempotent).  Eager steps and graph
        replays then share them: `self.lr` may be changed between steps, `self.step_count` keeps counting both."""
        if self._step_dev is None:
            dev = self.flat_param.device
            self._step_dev = torch.full((1,), self.step_count, device=dev, dtype=torch.int64)
            self._lr_dev = torch.full((1,), float(self.lr), device=dev, dtype=torch.float32)
            self._lr_on_device = float(self.lr)

    def _sync_lr(self) -> None:
        if self._lr_on_device != float(self.lr):
            self._lr_dev.fill_(float(self.lr))
            self._lr_on_device = float(self.lr)

    def capture(self, eeg: torch.Tensor, roi_series: torch.Tensor, conn: Optional[torch.Tensor] = None,
                warmup: int = 1) -> "GraphedStep":
        """-> the step on inputs of these shapes as ONE CUDA-graph launch (see GraphedStep)."""
        return GraphedStep(self, eeg, roi_series, conn, warmup)

    # -- end-to-end step from pinned host buffers --------------------------------------------
    def step_from_host(self, eeg_h: torch.Tensor, roi_h: torch.Tensor, conn_h: Optional[torch.Tensor] = None) -> float:
        """Host -> device copies of this step's inputs, the step, and the loss read back."""
        dev = self.flat_grad.device
        bufs = []
        for name, h in (("eeg", eeg_h), ("roi", roi_h), ("conn", conn_h)):
            if h is None:
                continue
            d = self._stage.get(name)
            if d is None or d.shape != h.shape:
                d = torch.empty(h.shape, device=dev, dtype=h.dtype)
                self._stage[name] = d
            d.copy_(h, non_blocking=True)
            bufs.append(d)
        return float(self.step(*bufs).item())


    def steps_from_host(self, host_batches) -> List[float]:
        """End-to-end epoch over pinned host batches [(eeg, roi, conn), ...]: the host->device copy of batch
        k+1 runs on a copy stream while batch k trains (two device buffer sets, event hand-off), and every
        step's loss is read back.  Returns the per-step losses."""
        dev = self.flat_grad.device
        main = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None) or torch.cuda.Stream(dev)
        self._copy_stream = copy
        sets = [dict(bufs=None, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        for st in sets:
            st["free"].record(main)

        def stage(st, batch):
            with torch.cuda.stream(copy):
                copy.wait_event(st["free"])  # the step that last read this buffer set has finished
                if st["bufs"] is None or any(d.shape != h.shape for d, h in zip(st["bufs"], batch)):
                    st["bufs"] = [torch.empty(h.shape, device=dev, dtype=h.dtype) for h in batch]
                for d, h in zip(st["bufs"], batch):
                    d.copy_(h, non_blocking=True)
                st["ready"].record(copy)

        batches = list(host_batches)
        losses: List[float] = []
        if not batches:
            return losses
        # every step's loss goes device -> pinned host asynchronously and is consumed one step later, so the
        # host keeps enqueueing step k+1 while step k runs (a blocking .item() per step would expose the
        # launch latency of ~300 kernels as GPU idle time)
        pinned = getattr(self, "_loss_pinned", None)
        if pinned is None or pinned.numel() < len(batches):
            pinned = torch.empty(max(len(batches), 16), dtype=torch.float32).pin_memory()
            self._loss_pinned = pinned
        done = [torch.cuda.Event() for _ in batches]
        stage(sets[0], batches[0])
        for k, _ in enumerate(batches):
            cur = sets[k % 2]
            if k + 1 < len(batches):
                stage(sets[(k + 1) % 2], batches[k + 1])
            main.wait_event(cur["ready"])
            loss = self.step(*cur["bufs"])
            cur["free"].record(main)
            pinned[k:k + 1].copy_(loss.reshape(1), non_blocking=True)
            done[k].record(main)
            if k > 0:
                done[k - 1].synchronize()
                losses.append(float(pinned[k - 1]))
        done[-1].synchronize()
        losses.append(float(pinned[len(batches) - 1]))
        return losses


class GraphedStep:
    """One `PairedTrainer.step` -- window gather, both encoders, InfoNCE, backward, clip_grad_norm_, AdamW: every launch
    of _test_bridge.py:775-788's recipe -- captured once in a CUDA graph and replayed as ONE launch.  The small
    configurations are bound by the ~200 C-ABI calls the host issues per step (BASELINE config 3, batch 256: 6.2 ms
    eager); a replay costs what the kernels cost.

    What a capture freezes, and how each piece still changes per replay:
      * inputs: static device buffers, refilled by `__call__` (device or pinned host tensors of the captured shapes);
      * dropout: every seed is a frozen kernel argument; the graph's first node advances the library's device-resident
        seed epoch, which every mask hash folds in (ops.seed_epoch_advance) -- fresh masks per replay, the same in the
        forward and the backward of one replay;
      * AdamW: step count and learning rate live in device memory (PairedTrainer.use_device_state): the count is
        incremented by a node of the graph, `trainer.lr` is written before the replay when it changed;
      * BatchNorm running statistics and num_batches_tracked are updated by captured device operations.
    The warm-up step(s) the capture needs (lazy initialisation must not happen inside a capture) run on a copy of the
    training state that is restored afterwards, so constructing a GraphedStep does not train.

    Data parallel (one process per GPU, every rank captures and replays in lockstep): NCCL's collectives and the
    symmetric-memory barrier are captured as they are; the SyncBN peer exchange publishes `seq + (seed epoch << 32)`, so a
    captured call stays distinct from replay to replay (csrc/peer_exchange.cu).  Its slot rotation is frozen with the
    capture: replays must follow each other without eager steps of the same trainer in between (checked), and a channel's
    calls per step must not be 1 modulo the slot count (checked at capture; they are 6 and 10)."""

    def __init__(self, trainer: "PairedTrainer", eeg: torch.Tensor, roi_series: torch.Tensor,
                 conn: Optional[torch.Tensor] = None, warmup: int = 1):
        dev = trainer.flat_param.device
        if dev.type != "cuda":
            raise ops._lib.XmodalError("GraphedStep needs a CUDA device")
        self.trainer = trainer
        ops.seed_epoch_init()
        trainer.use_device_state()
        self._in = [torch.empty(t.shape, device=dev, dtype=t.dtype).copy_(t, non_blocking=True) if t is not None else None
                    for t in (eeg, roi_series, conn)]
        args = [t for t in self._in if t is not None]
        cur = torch.cuda.current_stream(dev)
        self._stream = torch.cuda.Stream(dev)  # warm-up AND capture run here (the peer exchange keys its channels by stream)
        # -- warm-up on a snapshot of the training state
        model = trainer.model
        saved = {k: v.clone() for k, v in model.state_dict().items()}
        moments = (trainer.exp_avg.clone(), trainer.exp_avg_sq.clone())
        count, seeds = trainer.step_count, XF.seed_state()
        epoch = ops.seed_epoch_get()
        side = self._stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                trainer.step(*args)
        cur.wait_stream(side)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(saved[k])
            trainer.exp_avg.copy_(moments[0])
            trainer.exp_avg_sq.copy_(moments[1])
            trainer._step_dev.fill_(count)
        trainer.step_count = count
        XF.set_seed_state(seeds)  # (the warm-up drew host seeds; replays of this graph start from the same ones)
        ops.seed_epoch_set(epoch)
        trainer._sync_lr()
        # -- capture
        for p in trainer.params:
            p.grad = None
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        launches = ops.launch_count()
        before = XF.peer_exchange_counts()
        with torch.cuda.graph(self.graph, stream=self._stream, capture_error_mode="thread_local"):
            ops.seed_epoch_advance()
            self._loss = trainer.step(*args)
        self.launches_captured = ops.launch_count() - launches  # C-ABI calls inside the graph
        trainer.step_count = count  # the capture ran the host side of one step, not the device side
        self._peer_counts = XF.peer_exchange_counts()
        for ch, n in self._peer_counts.items():
            calls = n - before.get(ch, 0)
            if calls % XF._PeerReduce.SLOTS == 1:
                raise ops._lib.XmodalError(f"GraphedStep: {calls} peer exchanges per step on channel {ch}: a replay would reuse "
                                           "the slot of its predecessor's last call")

    def __call__(self, eeg: torch.Tensor, roi_series: torch.Tensor, conn: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Refill the static inputs (asynchronous copies on the current stream), replay, return the loss (device scalar)."""
        for dst, src in zip(self._in, (eeg, roi_series, conn)):
            if (dst is None) != (src is None):
                raise ValueError("GraphedStep: the captured step had a different set of inputs")
            if dst is not None and src is not dst:
                if src.shape != dst.shape:
                    raise ValueError(f"GraphedStep: input of shape {tuple(src.shape)}, captured {tuple(dst.shape)}")
                dst.copy_(src, non_blocking=True)
        return self.replay()

    def replay(self) -> torch.Tensor:
        """One step on whatever the static inputs (`self.inputs`) hold."""
        if XF.peer_exchange_counts() != self._peer_counts:
            raise ops._lib.XmodalError("GraphedStep: eager data-parallel steps ran since the capture / the last replay; the "
                                       "captured slot rotation of the peer exchange no longer lines up -- capture again")
        self.trainer._sync_lr()
        self.graph.replay()
        self.trainer.step_count += 1
        return self._loss.clone()

    @property
    def inputs(self):
        return self._in


class GraphedCallable:
    """Any training step built from this package's modules -- `fn(*static_inputs)` doing zero_grad -> forward -> backward
    -> clip -> optimizer.step, e.g. the recipes of run_training_lite.py:478-489 / run_fmri_v11.py:430-450 -- captured in a
    CUDA graph.  The graph's first node advances the seed epoch (fresh dropout masks per replay, see GraphedStep); the
    torch optimizers must be built with `capturable=True` (their step counters then live on the device).  `fn` runs
    `warmup` times before the capture; the state of `modules` and `optimizers` is restored afterwards (optimizer state the
    warm-up created is zeroed in place).  `fn` must not read device values back (no .item() / .tolist()).  `fn` may
    return a tensor (e.g. the loss): `__call__` returns a copy of it."""

    def __init__(self, fn, static_inputs, modules=(), optimizers=(), warmup: int = 1):
        import copy
        self.fn, self.inputs = fn, list(static_inputs)
        dev = self.inputs[0].device
        ops.seed_epoch_init()
        saved_m = [copy.deepcopy(m.state_dict()) for m in modules]
        saved_o = [copy.deepcopy(o.state_dict()) for o in optimizers]
        seeds, epoch = XF.seed_state(), ops.seed_epoch_get()
        cur, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                fn(*self.inputs)
        cur.wait_stream(side)
        for m, sd in zip(modules, saved_m):
            m.load_state_dict(sd)
        for o, sd in zip(optimizers, saved_o):
            if sd["state"]:
                o.load_state_dict(sd)
            else:  # state first created by the warm-up: keep the tensors (a capture must not allocate and re-zero them on
                # every replay) and reset them to the zeros torch's Adam-family optimizers start from
                for st in o.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        XF.set_seed_state(seeds)
        ops.seed_epoch_set(epoch)
        for m in modules:
            for p in m.parameters():
                p.grad = None
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        launches = ops.launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            ops.seed_epoch_advance()
            self._out = fn(*self.inputs)
        self.launches_captured = ops.launch_count() - launches

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._out.clone() if torch.is_tensor(self._out) else self._out


def train_bridge_epoch(model, loader, optimizer, criterion, device, grad_clip: float = 1.0) -> float:
    """_test_bridge.py:775-788 -- supervised bridge epoch (CE on logits), same recipe and return value."""
    model.train()
    total = 0.0
    for eeg, fmri, labels, _ in loader:
        eeg, fmri, labels = eeg.to(device), fmri.to(device), labels.to(device)
        optimizer.zero_grad()
        loss = criterion(model(eeg, fmri), labels)
        loss.backward()
        if grad_clip > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
        optimizer.step()
        total += loss.item()
    return total / max(len(loader), 1)
