#!/usr/bin/env python
"""Benchmark of the paired EEG/fMRI cross-modal training step (BASELINE.json metric:
paired EEG-fMRI train samples/s at 1/2/4/8 B200, with the roofline fraction of the dominant kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)

One "step" = one full training step of the composite model (EnhancedERPEncoder on 64 ch x 500 sample
windows + fMRIFusionNet on the mean/std and the 200 x 200 correlation matrix of 200 ROI x 100 TR series +
bridge projections + symmetric InfoNCE; backward; gradient all-reduce; clip_grad_norm_(1.0); AdamW) on a
per-GPU batch of 4096 synthetic paired samples (BASELINE config 4; weak scaling: global batch = 4096 x N
with global negatives).  The inputs of a step are the EEG windows and the ROI series; the 40 000-d
connectivity feature is derived from the ROI series on the device (`--conn host`
ships a precomputed connectivity matrix per sample instead, as round 1 did).  Prints ONE JSON line (rank 0)
which also carries: the whole-step roofline, the other BASELINE configs (1, 2, 3, 5) as `extras`, the same
step on the PyTorch library path on this GPU (`torch_eager_gpu`), and under torchrun a strong-scaling block
(global batch 4096 split over the ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "paired EEG-fMRI train samples/sec"
UNIT = "samples/s"
SHAPE = dict(eeg_channels=64, eeg_samples=500, n_roi=200, n_tr=100, conn_dim=40000)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md): ONE long-running
    `nvidia-smi -lms 200` process, started before the warm-up steps.  (Forking a new nvidia-smi every 200 ms put its
    driver initialisation inside the timed region: two of twelve runs of this bench lost ~90 ms of launches to it.)
    `mark()` notes where the timed region starts; only samples taken after it are summarised."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._proc, self._t, self._skip = index, [], None, None, 0

    def _run(self):
        for line in self._proc.stdout:
            line = line.strip()
            if line:
                self.rows.append([c.strip() for c in line.split(",")])

    def start(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                           "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
            # nvidia-smi takes 0.5 - 2 s to initialise NVML over all GPUs of the box and holds driver locks while it
            # does: wait for its first sample, so that neither the warm-up nor the timed region overlaps that start-up
            t0 = time.time()
            while not self.rows and time.time() - t0 < 10.0 and self._proc.poll() is None:
                time.sleep(0.02)
        except Exception:  # noqa: BLE001 - sampling is best effort
            self._proc = None
        return self

    def mark(self):
        self._skip = len(self.rows)

    def stop(self):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self._proc.kill()
            if self._t is not None:
                self._t.join(timeout=5)
        rows = self.rows[self._skip:]
        self.rows = rows if rows else self.rows[-2:]  # a very short timed region may fall between two samples

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ----------------------------------------------------------------------------- reference / CPU baseline arm
def _cpu_batch(requested: int) -> int:
    """The CPU arm runs the GPU arm's batch (4096, InfoNCE over 4096 negatives) when the host has the memory for
    the oracle's autograd tape (~9 MB per sample with the unfused attention weights); else a smaller sample."""
    if requested > 0:
        return requested
    try:
        import psutil
        free = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        free = 0
    for b in (4096, 2048, 1024, 512, 256):
        if free > b * 14e6 + 8e9:
            return b
    return 64


def cpu_reference_run(steps: int, warmup: int, sample_batch: int, max_seconds: float = 120.0, derive_conn: bool = True):
    """The reference's CPU path for the same step: oracle/paired_step.py (a functional restatement of
    the reference modules, pinned against golden vectors of the real classes, + the authored InfoNCE)
    on all host cores, on a bounded sample of the workload (`sample_batch` paired samples per step, as many
    steps as fit `max_seconds`)."""
    import torch

    from multimodal_eeg_fmri_b200 import synthetic
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel
    from oracle import paired_step as ops_

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    model = PairedBridgeModel(SHAPE["eeg_channels"], SHAPE["n_roi"], SHAPE["conn_dim"], 128, 64, 128, 0.0, 0.0, "v4")
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(sample_batch, SHAPE["eeg_channels"], SHAPE["eeg_samples"], SHAPE["n_roi"],
                                            SHAPE["n_tr"], SHAPE["conn_dim"], seed=42)
    if derive_conn:
        conn = None  # the oracle derives the connectivity from the ROI series, as the GPU arm does
    state = {}
    tw = time.perf_counter()
    for _ in range(warmup):
        ops_.paired_train_step(P, state, eeg, roi, conn, 0.07, "v4")
        if time.perf_counter() - tw > max_seconds / 3:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        ops_.paired_train_step(P, state, eeg, roi, conn, 0.07, "v4")
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": sample_batch * done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} steps of the same step on {sample_batch} paired samples (64ch x 500, 200 ROI x 100 TR, conn 40000 "
                      f"{'derived from the ROI series' if derive_conn else 'given'}; InfoNCE over {sample_batch} negatives), "
                      f"torch {torch.__version__} CPU fp32, dropout 0",
            "ms_per_step": 1e3 * dt / done, "steps": done, "batch": sample_batch}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = _cpu_batch(args.cpu_batch)
    r = cpu_reference_run(args.steps, min(args.warmup, 1), batch, max_seconds=100.0, derive_conn=args.conn == "device")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "paired bridge training step (BASELINE config 4), v4 ERP encoder, CPU batch %d "
                                   "(the GPU arm's per-GPU batch is 4096), connectivity %s" % (batch, args.conn), **SHAPE,
                       "per_gpu_batch": batch, "global_batch": batch},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- this repo's arm
# Whole-step roofline: 4.36 TFLOP (v4 encoder fwd + bwd) + 0.13 (fMRI net) + 0.013 (similarity)
# per 4096-sample step.  The north-star floor quotes the bf16 tensor rate; this implementation computes in tf32
# (the 1e-3 tolerance rules out bf16 operands: DESIGN.md section 2), whose rate is half.
STEP_TFLOP_PER_4096 = 4.36 + 0.13 + 0.013


def _timed(fn, steps, barrier, dev, world):
    import torch
    import torch.distributed as dist
    import gc
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Python's cyclic collector is kept out of the timed region: a generation-2 pass in the launch thread right after
    # the opening synchronize (empty GPU queue) stalled the first timed step by 100 - 200 ms in 3 of 20 runs
    # (tools/step_jitter.py, profiles/r2_step_jitter.txt); the step itself leaves no cyclic garbage (tools/gc_probe.py)
    gc.collect()
    gc.disable()
    try:
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
    finally:
        gc.enable()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def _extras(dev, rank, world, barrier, hbm):
    """The other BASELINE configs, each a short event-timed run (config 5 on every rank; 1, 2, 3 on one GPU)."""
    import torch

    from multimodal_eeg_fmri_b200 import eeg_data_utils as edu, ops, synthetic
    from multimodal_eeg_fmri_b200.modules import LabelSmoothingCrossEntropy, fMRIFusionNet
    from multimodal_eeg_fmri_b200.run_training_lite import ImprovedTriModalFusionNetLite
    from multimodal_eeg_fmri_b200.training import GraphedCallable, PairedBridgeModel, PairedTrainer

    out = {}
    # ---- config 5: band power of 128-channel 1 kHz recordings, win 1024 / hop 512, windows read in place
    C, win, hop, wpr, nrec = 128, 1024, 512, 64, 256
    n = win + (wpr - 1) * hop
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    bufs = [torch.randn(nrec, C, n, device=dev, generator=g) for _ in range(2)]  # 2 x 4.3 GB chunks (> L2), alternating
    chunks = 8
    for i in range(2):
        edu.band_power(bufs[i & 1], 1000.0, win, hop)
    ms = _timed(lambda it=iter(range(10 ** 9)): edu.band_power(bufs[next(it) & 1], 1000.0, win, hop), chunks, barrier, dev, world)
    nwin = chunks * nrec * wpr * world
    by = nwin * (C * win * 4 + C * 3 * 4)
    out["config5_bandpower"] = {"metric": "EEG band-power windows/sec", "value": round(nwin / (ms * 1e-3), 1), "unit": "windows/s",
                                "windows": nwin, "channels": C, "win": win, "hop": hop, "n_gpus": world, "ms_total": round(ms, 3),
                                "roofline": {"bound": "hbm", "achieved": round(by / (ms * 1e-3) / 1e9 / world, 1), "peak": hbm,
                                             "unit": "GB/s", "frac": round(by / (ms * 1e-3) / 1e9 / world / hbm, 4),
                                             "algorithmic_bytes_per_window": C * win * 4 + C * 3 * 4},
                                "note": "1M-window sweep = this rate x 1 048 576 windows; inputs resident in HBM, two alternating chunks"}
    del bufs
    if world > 1:
        return out  # configs 1-3 are single-GPU configurations

    def step_time(step, n_steps=20, warm=5):
        for _ in range(warm):
            step()
        return _timed(step, n_steps, barrier, dev, world) / n_steps

    # ---- config 3: paired InfoNCE bridge step, batch 256
    torch.manual_seed(42)
    m3 = PairedBridgeModel(SHAPE["eeg_channels"], SHAPE["n_roi"], SHAPE["conn_dim"], 128, 64, 128, 0.3, 0.4, "v4").to(dev).train()
    t3 = PairedTrainer(m3)
    eeg, roi, _ = (t.to(dev) for t in synthetic.paired_batch(256, SHAPE["eeg_channels"], SHAPE["eeg_samples"], SHAPE["n_roi"],
                                                             SHAPE["n_tr"], 16, seed=42))
    n0 = ops.launch_count()
    ms3 = step_time(lambda: t3.step(eeg, roi))
    out["config3_bridge_b256"] = {"metric": METRIC, "value": round(256 / (ms3 * 1e-3), 1), "unit": UNIT, "batch": 256,
                                  "ms_per_step": round(ms3, 3), "launches_per_step": (ops.launch_count() - n0) / 25}
    # the same step captured once and replayed as ONE CUDA-graph launch (PairedTrainer.capture: device-resident seed epoch,
    # AdamW step count and learning rate; replays equal eager steps bit for bit, tests/test_gpu_graphed_step.py)
    g3 = t3.capture(eeg, roi)
    ms3g = step_time(g3.replay)
    out["config3_bridge_b256"]["graphed"] = {"value": round(256 / (ms3g * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms3g, 3),
                                             "c_abi_calls_captured": g3.launches_captured, "graph_launches_per_step": 1}
    del g3, m3, t3
    # ---- config 1: run_training_lite step (tri-modal lite net, label-smoothing CE, AdamW 5e-5 / 0.01, clip 1.0), batch 32
    torch.manual_seed(42)
    m1 = ImprovedTriModalFusionNetLite(64, 64, 6048).to(dev).train()
    crit = LabelSmoothingCrossEntropy(0.1)
    opt1 = torch.optim.AdamW(m1.parameters(), lr=5e-5, weight_decay=0.01, fused=True, capturable=True)
    erp, pw, cn = torch.randn(32, 64, 500, device=dev), torch.randn(32, 64, 500, device=dev), torch.randn(32, 6048, device=dev)
    y = torch.randint(0, 2, (32,), device=dev)

    def step1():
        opt1.zero_grad(set_to_none=True)
        crit(m1(pw, erp, cn), y).backward()
        torch.nn.utils.clip_grad_norm_(m1.parameters(), 1.0)
        opt1.step()
    ms1 = step_time(step1)
    out["config1_lite_b32"] = {"metric": "tri-modal lite train samples/sec", "value": round(32 / (ms1 * 1e-3), 1), "unit": UNIT,
                               "batch": 32, "ms_per_step": round(ms1, 3)}
    # (capturable because the wrapper's fusion weights stay on the device until somebody reads them, modules.DeviceFloats;
    #  the reference reads five scalars back in every forward, crossmodal_v4_enhancements.py:803-806)
    g1 = GraphedCallable(lambda *_: step1(), [pw, erp, cn, y], [m1], [opt1])
    ms1g = step_time(g1.graph.replay)
    out["config1_lite_b32"]["graphed"] = {"value": round(32 / (ms1g * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms1g, 3),
                                          "c_abi_calls_captured": g1.launches_captured}
    del g1
    # ---- config 2: run_fmri_v11 step (fMRIFusionNet 400 / 40 000, weighted CE, AdamW 1e-4, clip 1.0), batch 64
    torch.manual_seed(42)
    m2 = fMRIFusionNet(400, 40000).to(dev).train()
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-4, weight_decay=1e-4, fused=True, capturable=True)
    act, conn2, y2 = torch.randn(64, 400, device=dev), torch.randn(64, 40000, device=dev), torch.randint(0, 2, (64,), device=dev)

    def step2():
        opt2.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(m2(act, conn2), y2).backward()
        torch.nn.utils.clip_grad_norm_(m2.parameters(), 1.0)
        opt2.step()
    ms2 = step_time(step2)
    out["config2_fmri_b64"] = {"metric": "fMRI ROI encoder train samples/sec", "value": round(64 / (ms2 * 1e-3), 1), "unit": UNIT,
                               "batch": 64, "ms_per_step": round(ms2, 3)}
    g2 = GraphedCallable(lambda *_: step2(), [act, conn2, y2], [m2], [opt2])
    ms2g = step_time(g2.graph.replay)
    out["config2_fmri_b64"]["graphed"] = {"value": round(64 / (ms2g * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms2g, 3),
                                          "c_abi_calls_captured": g2.launches_captured}
    del g2
    return out


def _torch_eager_gpu(dev, batch, derive_conn, barrier):
    """SURVEY.md section 2's bar: the SAME step on the PyTorch library path of this GPU (cuDNN convolutions, cuBLAS
    matmuls, torch's unfused attention, foreach optimizer) -- the functional restatement of the reference modules
    (oracle/) run on cuda tensors.  A baseline leg like cpu_baseline: measured, not shipped."""
    import torch

    from multimodal_eeg_fmri_b200 import synthetic
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel
    from oracle import paired_step as ops_

    res = {}
    torch.manual_seed(42)
    model = PairedBridgeModel(SHAPE["eeg_channels"], SHAPE["n_roi"], SHAPE["conn_dim"], 128, 64, 128, 0.0, 0.0, "v4")
    P0 = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
    eeg, roi, conn = (t.to(dev) for t in synthetic.paired_batch(batch, SHAPE["eeg_channels"], SHAPE["eeg_samples"], SHAPE["n_roi"],
                                                                SHAPE["n_tr"], SHAPE["conn_dim"] if not derive_conn else 16, seed=42))
    if derive_conn:
        conn = None
    for name, tf32 in (("tf32", True), ("fp32", False)):
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
        try:
            P, state = {k: v.clone() for k, v in P0.items()}, {}
            for _ in range(2):
                ops_.paired_train_step(P, state, eeg, roi, conn, 0.07, "v4")
            ms = _timed(lambda: ops_.paired_train_step(P, state, eeg, roi, conn, 0.07, "v4"), 3, barrier, dev, 1) / 3
            res[name] = {"value": round(batch / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 3)}
        except torch.cuda.OutOfMemoryError:
            res[name] = {"unavailable": f"out of device memory at batch {batch}"}
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
            del P, state
            torch.cuda.empty_cache()
    res["batch"] = batch
    res["what"] = ("oracle/paired_step.py (functional restatement of the reference modules) on cuda tensors, dropout 0: cuDNN / "
                   "cuBLAS / unfused attention, allow_tf32 on | off; same inputs and batch as the main line")
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    from multimodal_eeg_fmri_b200 import _lib, ops, synthetic
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer, init_distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    _lib.lib()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _bind_to_gpu_numa_node(local)
    ctx = init_distributed("nccl")
    dev = torch.device("cuda", local)
    B = args.batch
    derive = args.conn == "device"

    torch.manual_seed(42)  # identical initial weights on every rank
    model = PairedBridgeModel(SHAPE["eeg_channels"], SHAPE["n_roi"], SHAPE["conn_dim"], 128, 64, 128, 0.3, 0.4, args.encoder)
    model = model.to(dev).train()
    trainer = PairedTrainer(model, lr=1e-4, weight_decay=1e-4, grad_clip=1.0)

    def make_inputs(batch, offset):
        ts = synthetic.paired_batch(batch, SHAPE["eeg_channels"], SHAPE["eeg_samples"], SHAPE["n_roi"], SHAPE["n_tr"],
                                    SHAPE["conn_dim"] if not derive else 16, seed=42, offset=offset)
        return [t.pin_memory() for t in (ts[:2] if derive else ts)]

    # this rank's rows of the global synthetic batch, in pinned host memory (the e2e source) and in HBM
    host = make_inputs(B, rank * B)
    devt = [t.to(dev, non_blocking=True) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local).start()  # started (and initialised) before the warm-up: its start-up stays outside the timed region
    # W warm-up steps as asked, plus 5 settling steps of our own (all untimed): the caching allocator of the two streams
    # reaches its steady state and every kernel has been loaded before the timed region opens
    for _ in range(max(args.warmup, 3) + 5):
        trainer.step(*devt)
    barrier()

    # ---- timed region 1: device-resident inputs
    n0 = ops.launch_count()
    clocks.mark()
    ms_total = _timed(lambda: trainer.step(*devt), args.steps, barrier, dev, world)
    clocks.stop()
    launches = ops.launch_count() - n0
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    # ---- instrumented pass (NOT part of `value`): the same steps with the two model branches serialised on one stream
    # and CUDA events around every C-ABI call, so that each entry point's time is that of its kernels running alone
    overlap = model.overlap_branches
    model.overlap_branches = False
    tl_steps = min(args.steps, 5)
    trainer.step(*devt)
    ops.start_timeline()
    ms_serial = _timed(lambda: trainer.step(*devt), tl_steps, barrier, dev, world) / tl_steps
    timeline = ops.stop_timeline()
    model.overlap_branches = overlap

    # ---- timed region 2: end to end from pinned host buffers (H2D of the inputs + loss read back, every step)
    # (the public end-to-end call: copies batch k+1 from pinned host memory while batch k trains; every step's
    #  inputs cross PCIe inside the timed region and every step's loss is read back)
    if args.e2e_steps <= 0:
        # the first batch's host->device copy cannot hide behind a previous step (15 ms at 852 MB): 30 steps keep that
        # pipeline fill at ~0.5 ms per step instead of 1.5 ms with 10 (an epoch of the reference has hundreds of steps)
        args.e2e_steps = max(args.steps, 30)
    trainer.steps_from_host([host, host])
    ms_e2e = _timed(lambda: trainer.steps_from_host([host] * args.e2e_steps), 1, barrier, dev, world)
    e2e_value = world * B * args.e2e_steps / (ms_e2e * 1e-3)
    loss_t = trainer.step(*devt).clone()
    if world > 1:
        dist.all_reduce(loss_t)  # global loss = sum of the per-rank shares
    loss = float(loss_t)

    # ---- strong scaling (SURVEY.md section 8e as written): the GLOBAL batch stays 4096, rank r holds 4096 / N rows
    strong = strong_inputs = None
    if world > 1 and B % world == 0:
        Bs = B // world
        sh = [t.to(dev) for t in make_inputs(Bs, rank * Bs)]
        for _ in range(3):
            trainer.step(*sh)
        ms_s = _timed(lambda: trainer.step(*sh), args.steps, barrier, dev, world) / args.steps
        strong_inputs = sh
        strong = {"global_batch": B, "per_gpu_batch": Bs, "ms_per_step": round(ms_s, 3), "value": round(B / (ms_s * 1e-3), 1),
                  "unit": UNIT, "scaling": "strong", "note": "same global batch as the 1-GPU line; in-GEMM peer reads for per-GPU batch <= 512"}
    peak_mem = torch.cuda.max_memory_allocated()
    torch.cuda.empty_cache()

    hbm, tc_burst, tc_sus, src = _peaks()
    extras = _extras(dev, rank, world, barrier, hbm) if not args.no_extras else {}
    kernels = {}
    for name, (n, ms, fl, by) in sorted(timeline.items(), key=lambda kv: -kv[1][1]):
        kernels[name] = {"calls_per_step": n / tl_steps, "ms_per_step": round(ms / tl_steps, 4),
                         "tflops": round(fl / (ms * 1e-3) / 1e12, 2) if ms > 0 else None,
                         "gbs": round(by / (ms * 1e-3) / 1e9, 1) if ms > 0 else None}
    ours_ms = sum(v[1] for v in timeline.values()) / tl_steps
    # Dominant kernel: gemm_tf32_kernel<EPI_ROWMAJOR> (tcgen05 TF32 engine) as launched by the linear / conv
    # fwd, dgrad and wgrad entry points.  At d_model = 128 its arithmetic intensity (50-100 FLOP/B) is at or
    # below the tf32 ridge, so the bound is HBM: achieved = algorithmic operand bytes / CUDA-event time.
    names = ("xm_linear_fwd_f32", "xm_linear_fwd_stacked3_f32", "xm_linear_dgrad_f32", "xm_linear_wgrad_f32",
             "xm_infonce_dgrad_f32", "xm_linear_dgrad_peers_f32", "xm_conv1d_fwd_f32", "xm_conv1d_fwd_stats_f32",
             "xm_conv1d_dgrad_f32", "xm_conv1d_wgrad_f32")
    gk = [timeline[k] for k in names if k in timeline]
    g_ms, g_fl, g_by, g_n = sum(v[1] for v in gk), sum(v[2] for v in gk), sum(v[3] for v in gk), sum(v[0] for v in gk)
    ach = g_by / (g_ms * 1e-3) / 1e9 if g_ms > 0 else 0.0
    traffic = None
    tj = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tj):
        tr_ = json.load(open(tj))
        traffic = {"dram_bytes_per_launch": tr_["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": tr_["algorithmic_bytes_per_launch"],
                   "case": tr_["case"], "source": tr_["source"]}
    roofline = {"bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4),
                "traffic": traffic, "kernel": "gemm_tf32_kernel<EPI_ROWMAJOR> via xm_{linear,conv1d}_{fwd,dgrad,wgrad}_f32, xm_conv1d_fwd_stats_f32, xm_linear_fwd_stacked3_f32",
                "launches_per_step": g_n / tl_steps, "ms_per_step": round(g_ms / tl_steps, 4),
                "share_of_step": round(g_ms / tl_steps / ms_serial, 4),
                "peak_source": f"{src}: hbm_gbs (copy bandwidth)",
                "tensor": {"achieved_tflops": round(g_fl / (g_ms * 1e-3) / 1e12, 1), "peak_tflops": round(tc_sus / 2.0, 1),
                           "note": "kind::tf32 peak = measured bf16_tflops_sustained / 2"}}
    step_tflop = STEP_TFLOP_PER_4096 * B / 4096.0
    floor_bf16, floor_tf32 = step_tflop / tc_sus * 1e3, step_tflop / (tc_sus / 2.0) * 1e3
    step_roofline = {"bound": "tensor", "work_tflop_per_step": round(step_tflop, 3), "floor_ms_bf16_rate": round(floor_bf16, 3),
                     "floor_ms_tf32_rate": round(floor_tf32, 3), "frac_of_bf16_floor": round(floor_bf16 / ms_step, 4),
                     "frac_of_tf32_floor": round(floor_tf32 / ms_step, 4),
                     "note": "SURVEY 8d work per step over the measured sustained tensor rate; the path computes in tf32 (half the bf16 rate)"}
    def make_line(cpu=None, eager=None, graphed=None):
      line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
              "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
              "data": "synthetic",
              "config": {"workload": f"paired bridge training step (BASELINE config 4), per-GPU batch {B}, global batch {B * world}, "
                                     f"{args.encoder} ERP encoder + fMRIFusionNet + bridge projections + symmetric InfoNCE (global negatives), "
                                     "backward, grad all-reduce, clip 1.0, AdamW",
                         **SHAPE, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                         "connectivity": ("derived on the device from the ROI series (per-sample corrcoef, xm_roi_corrcoef_f32)" if derive
                                          else "precomputed per sample, shipped from the host"),
                         "l2": f"inputs {h2d / 1e6:.0f} MB per step and >10 GB of activations per step, far larger than the 126 MB L2 (no flush needed)",
                         "final_loss": round(loss, 5)},
              "clocks": clocks.summary(),
              "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                      "ms_per_step": round(ms_e2e / args.e2e_steps, 3), "steps": args.e2e_steps},
              "gpu_launches": launches,
              "roofline": roofline,
              "step_roofline": step_roofline,
              "peak_device_memory_gb": round(peak_mem / 2 ** 30, 2),
              "branch_overlap": bool(overlap),
              "instrumented_pass": {"ms_per_step": round(ms_serial, 3), "steps": tl_steps, "own_kernels_ms_per_step": round(ours_ms, 3),
                                    "torch_ops_ms_per_step": round(ms_serial - ours_ms, 3),
                                    "note": "branches serialised, CUDA events around every C-ABI call: source of `kernels` and `roofline`"},
              "kernels": kernels}
      if strong is not None:
          line["strong_scaling"] = strong
      if extras:
          line["extras"] = extras
      if eager is not None:
          line["torch_eager_gpu"] = eager
      if cpu is not None:
          line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
      if graphed is not None:
          line["graphed_step"] = graphed
      return line

    # ---- the same step as ONE CUDA-graph launch (PairedTrainer.capture; replays equal eager steps bit for bit, also data
    # parallel: tests/test_gpu_graphed_step.py, tests/dp_graph_check.py).  `value` above stays the eager call; this block is
    # measured last on every rank, under a watchdog: should a capture or replay ever hang, rank 0 still prints the line.
    import threading

    def on_timeout():
        if rank == 0:
            print(json.dumps(make_line(graphed={"error": "timed out (watchdog)"})), flush=True)
        os._exit(0)

    graphed = None
    if not args.no_graph:
        watchdog = threading.Timer(240.0, on_timeout)
        watchdog.daemon = True
        watchdog.start()
        try:
            g = trainer.capture(*devt)
            for _ in range(3):
                g.replay()
            ms_g = _timed(g.replay, args.steps, barrier, dev, world) / args.steps
            graphed = {"ms_per_step": round(ms_g, 3), "value": round(world * B / (ms_g * 1e-3), 1), "unit": UNIT, "scaling": "weak",
                       "c_abi_calls_captured": g.launches_captured, "graph_launches_per_step": 1,
                       "note": "device-resident inputs, same per-GPU batch as `value`"}
            del g
            if strong_inputs is not None:
                gs = trainer.capture(*strong_inputs)
                for _ in range(3):
                    gs.replay()
                ms_gs = _timed(gs.replay, args.steps, barrier, dev, world) / args.steps
                graphed["strong_scaling"] = {"global_batch": B, "per_gpu_batch": B // world, "ms_per_step": round(ms_gs, 3),
                                             "value": round(B / (ms_gs * 1e-3), 1), "unit": UNIT}
                del gs
        except Exception as exc:  # noqa: BLE001 - an extra: never costs the line
            graphed = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        watchdog.cancel()
    del devt, strong_inputs
    torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = eager = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_reference_run(args.cpu_steps, 1, _cpu_batch(args.cpu_batch), max_seconds=30.0, derive_conn=derive)
    if world == 1 and not args.no_eager:
        free = torch.cuda.mem_get_info(dev)[0]
        eager = _torch_eager_gpu(dev, B if free > 150e9 else B // 4, derive, barrier)
    print(json.dumps(make_line(cpu, eager, graphed)), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _bind_to_gpu_numa_node(local: int) -> None:
    """Pin this rank's host threads (and so its pinned staging buffers, first-touch) to the CPU cores next to its GPU:
    with 8 ranks pulling 0.85 GB per step each over PCIe, buffers on the far socket halve the copy rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = (os.cpu_count() or 1 + 63) // 64 + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001 - best effort: NVML or the affinity call may be unavailable in the container
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="per-GPU batch (paired samples)")
    ap.add_argument("--encoder", default="v4", choices=["v4", "lite"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end region (0: max(--steps, 30))")
    ap.add_argument("--cpu-batch", type=int, default=0, help="paired samples per CPU-baseline step (0: the GPU arm's batch if host memory allows)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch_eager_gpu baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (1, 2, 3, 5)")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay block (`graphed_step`)")
    ap.add_argument("--conn", default="device", choices=["device", "host"],
                    help="connectivity features: derived on the device from the ROI series | precomputed, shipped from the host")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
