#!/usr/bin/env python
"""Benchmark of the paired EEG/fMRI cross-modal training step (BASELINE.json metric:
paired EEG-fMRI train samples/s at 1/2/4/8 B200, with the roofline fraction of the dominant kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)

One "step" = one full training step of the composite model (EnhancedERPEncoder on 64 ch x 500 sample
windows + fMRIFusionNet on 200 ROI x 100 TR series / 40 000-d connectivity + bridge projections +
symmetric InfoNCE; backward; gradient all-reduce; clip_grad_norm_(1.0); AdamW) on a per-GPU batch of
4096 synthetic paired samples (BASELINE config 4; weak scaling: global batch = 4096 x N with global
negatives).  Prints ONE JSON line (rank 0).
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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:  # noqa: BLE001 - sampling is best effort
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ----------------------------------------------------------------------------- reference / CPU baseline arm
def cpu_reference_run(steps: int, warmup: int, sample_batch: int, max_seconds: float = 120.0):
    """The reference's CPU path for the same step: oracle/paired_step.py (a functional restatement of
    the reference modules, pinned against golden vectors of the real classes, + the authored InfoNCE)
    on all host cores, on a bounded sample of the workload (`sample_batch` paired samples per step)."""
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
    state = {}
    for _ in range(warmup):
        ops_.paired_train_step(P, state, eeg, roi, conn, 0.07, "v4")
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        ops_.paired_train_step(P, state, eeg, roi, conn, 0.07, "v4")
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": sample_batch * done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} steps of the same step on {sample_batch} paired samples (64ch x 500, 200 ROI x 100 TR, conn 40000; "
                      f"InfoNCE over {sample_batch} negatives), torch {torch.__version__} CPU fp32, dropout 0",
            "ms_per_step": 1e3 * dt / done, "steps": done}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, args.cpu_batch)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "paired bridge training step, v4 ERP encoder, CPU sample batch %d" % args.cpu_batch, **SHAPE},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- this repo's arm
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
    ctx = init_distributed("nccl")
    dev = torch.device("cuda", local)
    B = args.batch

    torch.manual_seed(42)  # identical initial weights on every rank
    model = PairedBridgeModel(SHAPE["eeg_channels"], SHAPE["n_roi"], SHAPE["conn_dim"], 128, 64, 128, 0.3, 0.4, args.encoder)
    model = model.to(dev).train()
    trainer = PairedTrainer(model, lr=1e-4, weight_decay=1e-4, grad_clip=1.0)

    # this rank's rows of the global synthetic batch, in pinned host memory (the e2e source) and in HBM
    host = [t.pin_memory() for t in synthetic.paired_batch(B, SHAPE["eeg_channels"], SHAPE["eeg_samples"], SHAPE["n_roi"],
                                                            SHAPE["n_tr"], SHAPE["conn_dim"], seed=42, offset=rank * B)]
    devt = [t.to(dev, non_blocking=True) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        trainer.step(*devt)
    barrier()

    # ---- timed region 1: device-resident inputs, per-entry-point CUDA-event timeline on the launching stream
    n0 = ops.launch_count()
    ops.start_timeline()
    with ClockSampler(local) as clocks:
        ms_total = timed(lambda: trainer.step(*devt), args.steps)
    timeline = ops.stop_timeline()
    launches = ops.launch_count() - n0
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- timed region 2: end to end from pinned host buffers (H2D of the inputs + loss read back, every step)
    # (the public end-to-end call: copies batch k+1 from pinned host memory while batch k trains; every step's
    #  inputs cross PCIe inside the timed region and every step's loss is read back)
    trainer.steps_from_host([host, host])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    trainer.steps_from_host([host] * args.e2e_steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t)
    e2e_value = world * B * args.e2e_steps / (ms_e2e * 1e-3)
    loss_t = trainer.step(*devt).clone()
    if world > 1:
        dist.all_reduce(loss_t)  # global loss = sum of the per-rank shares
    loss = float(loss_t)

    if rank != 0:
        return
    hbm, tc_burst, tc_sus, src = _peaks()
    kernels = {}
    for name, (n, ms, fl, by) in sorted(timeline.items(), key=lambda kv: -kv[1][1]):
        kernels[name] = {"calls_per_step": n / args.steps, "ms_per_step": round(ms / args.steps, 4),
                         "tflops": round(fl / (ms * 1e-3) / 1e12, 2) if ms > 0 else None,
                         "gbs": round(by / (ms * 1e-3) / 1e9, 1) if ms > 0 else None}
    ours_ms = sum(v[1] for v in timeline.values()) / args.steps
    # Dominant kernel: gemm_tf32_kernel<EPI_ROWMAJOR> (tcgen05 TF32 engine) as launched by the linear / conv
    # fwd, dgrad and wgrad entry points.  At d_model = 128 its arithmetic intensity (50-100 FLOP/B) is at or
    # below the tf32 ridge, so the bound is HBM: achieved = algorithmic operand bytes / CUDA-event time.
    names = ("xm_linear_fwd_f32", "xm_linear_fwd_stacked3_f32", "xm_linear_dgrad_f32", "xm_linear_wgrad_f32",
             "xm_infonce_dgrad_f32", "xm_linear_dgrad_peers_f32", "xm_conv1d_fwd_f32", "xm_conv1d_dgrad_f32",
             "xm_conv1d_wgrad_f32")
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
                "traffic": traffic, "kernel": "gemm_tf32_kernel<EPI_ROWMAJOR> via xm_{linear,conv1d}_{fwd,dgrad,wgrad}_f32, xm_linear_fwd_stacked3_f32, xm_infonce_dgrad_f32",
                "launches_per_step": g_n / args.steps, "ms_per_step": round(g_ms / args.steps, 4),
                "share_of_step": round(g_ms / args.steps / ms_step, 4),
                "peak_source": f"{src}: hbm_gbs (copy bandwidth)",
                "tensor": {"achieved_tflops": round(g_fl / (g_ms * 1e-3) / 1e12, 1), "peak_tflops": round(tc_sus / 2.0, 1),
                           "note": "kind::tf32 peak = measured bf16_tflops_sustained / 2"}}
    cpu = cpu_reference_run(args.cpu_steps, 1, args.cpu_batch, max_seconds=30.0) if world == 1 and not args.no_cpu else None
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
            "data": "synthetic",
            "config": {"workload": f"paired bridge training step (BASELINE config 4), per-GPU batch {B}, global batch {B * world}, "
                                   f"{args.encoder} ERP encoder + fMRIFusionNet + bridge projections + symmetric InfoNCE (global negatives), "
                                   "backward, grad all-reduce, clip 1.0, AdamW",
                       **SHAPE, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": f"inputs {h2d / 1e6:.0f} MB per step and >10 GB of activations per step, far larger than the 126 MB L2 (no flush needed)",
                       "final_loss": round(loss, 5)},
            "clocks": clocks.summary(),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / args.e2e_steps, 3), "steps": args.e2e_steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "peak_device_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
            "own_kernels_ms_per_step": round(ours_ms, 3),
            "torch_ops_ms_per_step": round(ms_step - ours_ms, 3),
            "kernels": kernels}
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="per-GPU batch (paired samples)")
    ap.add_argument("--encoder", default="v4", choices=["v4", "lite"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-batch", type=int, default=64, help="paired samples per CPU-baseline step (bounded sample)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
